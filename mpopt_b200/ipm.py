"""A small primal-dual interior-point NLP solver that CONSUMES the evaluators (host code, not part of the hot path).

The reference hands its NLP to IPOPT through ``ca.nlpsol`` (/root/reference/mpopt/mpopt.py:757, :804); IPOPT is not
installable in this image.  SciPy's SLSQP / trust-constr stall well above 1e-6 on the stiffer transcriptions (the
hyper-sensitive problem at K=5, p=50), which is not enough to compare an optimum with the 17-digit objectives the
reference's notebooks store.  This module is the stand-in: a line-search barrier method in the spirit of IPOPT's
algorithm (Waechter & Biegler 2006) -- slack variables for two-sided constraint rows, fixed variables removed,
fraction-to-the-boundary rule, inertia-corrected KKT systems (dense LDL^T, LAPACK dsytrf), monotone barrier update,
l1 merit function with Armijo backtracking -- driven by exact first and second derivatives:

    f(x), grad_f(x), g(x), jac_g(x) (scipy.sparse), hess_l(x, lam_f, lam_g) (lower triangle, scipy.sparse)

i.e. exactly the five callbacks IPOPT takes.  Sized for the problems the reference's docs and tests solve
(n + m up to a few thousand: dense KKT matrices).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg.lapack as lapack
import scipy.sparse as sp


class IpmResult(dict):
    __getattr__ = dict.get


def _inertia(ldu, ipiv):
    """(n_pos, n_neg, n_zero) of the block-diagonal factor of dsytrf (lower)."""
    n = ldu.shape[0]
    pos = neg = zero = 0
    k = 0
    while k < n:
        if ipiv[k] > 0:
            d = ldu[k, k]
            if d > 0:
                pos += 1
            elif d < 0:
                neg += 1
            else:
                zero += 1
            k += 1
        else:  # 2 x 2 pivot block: symmetric [[a, b], [b, c]]
            a, b, c = ldu[k, k], ldu[k + 1, k], ldu[k + 1, k + 1]
            tr, det = a + c, a * c - b * b
            if det < 0:
                pos += 1
                neg += 1
            elif det > 0:
                if tr > 0:
                    pos += 2
                else:
                    neg += 2
            else:
                zero += 1
                pos += tr > 0
                neg += tr < 0
            k += 2
    return pos, neg, zero


def solve_nlp(*args, **kw):
    """See :func:`_solve_nlp` (same arguments)."""
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):  # infinite bounds take part in masked arithmetic
        return _solve_nlp(*args, **kw)


def _solve_nlp(f, grad_f, g, jac_g, hess_l, x0, lbx, ubx, lbg, ubg, tol=1e-9, max_iter=500, mu0=0.1, lam_g0=None,
              verbose=False, acceptable_tol=1e-6, acceptable_iter=15, bound_push=1e-2):
    """Minimise f(x) s.t. lbg <= g(x) <= ubg, lbx <= x <= ubx.  Returns IpmResult(x, f, g, lam_g, success, iter, err).

    Warm start (IPOPT's warm_start_init_point): pass the previous ``lam_g`` as ``lam_g0`` together with a small ``mu0``
    and ``bound_push`` (e.g. 1e-7 / 1e-9) so that the starting point is not pushed away from its active bounds."""
    x = np.array(x0, dtype=float)
    lbx, ubx, lbg, ubg = (np.asarray(v, dtype=float) for v in (lbx, ubx, lbg, ubg))
    n_all, m_all = x.size, lbg.size
    fixed = lbx == ubx
    x[fixed] = lbx[fixed]
    free = np.flatnonzero(~fixed)
    eq = lbg == ubg
    ie = np.flatnonzero(~eq)
    n, ns, m = free.size, ie.size, m_all
    ny = n + ns
    L = np.concatenate([lbx[free], lbg[ie]])
    U = np.concatenate([ubx[free], ubg[ie]])
    hasL, hasU = np.isfinite(L), np.isfinite(U)
    target = np.where(eq, lbg, 0.0)  # c = g - target (equalities) | g - s (inequalities)

    def push(y):
        y = y.copy()
        k1 = k2 = float(bound_push)
        pl = np.where(hasL & hasU, np.minimum(k1 * np.maximum(1.0, np.abs(L)), k2 * (U - L)), k1 * np.maximum(1.0, np.abs(L)))
        pu = np.where(hasL & hasU, np.minimum(k1 * np.maximum(1.0, np.abs(U)), k2 * (U - L)), k1 * np.maximum(1.0, np.abs(U)))
        lo = np.where(hasL, L + pl, -np.inf)
        hi = np.where(hasU, U - pu, np.inf)
        return np.minimum(np.maximum(y, lo), hi)

    def unpack(y):
        xx = x.copy()
        xx[free] = y[:n]
        return xx, y[n:]

    y = push(np.concatenate([x[free], g(x)[ie]]))
    if lam_g0 is None:
        zL, zU = np.where(hasL, 1.0, 0.0), np.where(hasU, 1.0, 0.0)
    else:  # warm start: duals on the central path of the starting barrier parameter
        zL = np.where(hasL, mu0 / np.where(hasL, y - L, 1.0), 0.0)
        zU = np.where(hasU, mu0 / np.where(hasU, U - y, 1.0), 0.0)
    lam = np.zeros(m) if lam_g0 is None else np.array(lam_g0, dtype=float)
    mu = float(mu0)
    filt, filt_mu, th_max, th_min = [], None, np.inf, 0.0
    kappa_eps, kappa_mu, theta_mu, tau_min = 10.0, 0.2, 1.5, 0.99
    delta_w_last = 0.0

    def evaluate(y):
        xx, s = unpack(y)
        gv = g(xx)
        c = gv - target
        c[ie] -= s
        return xx, s, float(f(xx)), gv, c

    def barrier(y, fv, mu):
        return fv - mu * (np.log(y[hasL] - L[hasL]).sum() + np.log(U[hasU] - y[hasU]).sum())

    def jac_y(xx):
        J = sp.csr_matrix(jac_g(xx))
        A = np.zeros((m, ny))
        A[:, :n] = J[:, free].toarray()
        A[ie, n + np.arange(ns)] = -1.0
        return A

    def error(gy, A, lam, c, y, zL, zU, mu):
        dual = gy + A.T @ lam - zL + zU
        sd = max(100.0, (np.abs(lam).sum() + zL.sum() + zU.sum()) / max(1, m + 2 * ny)) / 100.0
        sc = max(100.0, (zL.sum() + zU.sum()) / max(1, 2 * ny)) / 100.0
        compL = ((y - L) * zL - mu)[hasL]
        compU = ((U - y) * zU - mu)[hasU]
        comp = max(np.abs(compL).max(initial=0.0), np.abs(compU).max(initial=0.0))
        return max(np.abs(dual).max(initial=0.0) / sd, np.abs(c).max(initial=0.0), comp / sc)

    xx, s, fv, gv, c = evaluate(y)
    it, ok, acc_count = 0, False, 0
    err0 = np.inf
    for it in range(max_iter + 1):
        gx = np.asarray(grad_f(xx), dtype=float)
        gy = np.concatenate([gx[free], np.zeros(ns)])
        A = jac_y(xx)
        if it == 0 and lam_g0 is None:  # least-squares multipliers
            rhs = -(gy - zL + zU)
            try:
                lam = np.linalg.lstsq(A.T, rhs, rcond=None)[0]
                if np.abs(lam).max(initial=0.0) > 1e3:
                    lam[:] = 0.0
            except np.linalg.LinAlgError:
                lam[:] = 0.0
        err0 = error(gy, A, lam, c, y, zL, zU, 0.0)
        if verbose:
            print(f"{it:4d} f={fv:+.12e} err={err0:.2e} mu={mu:.1e} |c|={np.abs(c).max(initial=0):.1e} dw={delta_w_last:.1e} "
                  f"a_pr={locals().get('a_pr', 0):.2e} alpha={locals().get('alpha', 0):.2e} a_du={locals().get('a_du', 0):.2e}")
        if err0 <= tol:
            ok = True
            break
        acc_count = acc_count + 1 if err0 <= acceptable_tol else 0
        if acc_count >= acceptable_iter:
            ok = True
            break
        if it == max_iter:
            break
        while mu > tol / 10.0 and error(gy, A, lam, c, y, zL, zU, mu) <= kappa_eps * mu:
            mu = max(tol / 10.0, min(kappa_mu * mu, mu ** theta_mu))
        tau = max(tau_min, 1.0 - mu)
        # ---- KKT system
        Hl = sp.csr_matrix(hess_l(xx, 1.0, lam))
        Hd = Hl.toarray()
        Hd = Hd + np.tril(Hd, -1).T
        W = np.zeros((ny, ny))
        W[:n, :n] = Hd[np.ix_(free, free)]
        dL = np.where(hasL, y - L, 1.0)
        dU = np.where(hasU, U - y, 1.0)
        Sig = np.where(hasL, zL / dL, 0.0) + np.where(hasU, zU / dU, 0.0)
        r1 = gy + A.T @ lam - np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0)
        rhs = -np.concatenate([r1, c])
        Kbase = np.zeros((ny + m, ny + m))
        Kbase[:ny, :ny] = W + np.diag(Sig)
        Kbase[ny:, :ny] = A
        Kbase[:ny, ny:] = A.T
        dw, dc = 0.0, 0.0
        sol = None
        for attempt in range(40):
            Kt = Kbase.copy()
            if dw:
                Kt[np.arange(ny), np.arange(ny)] += dw
            if dc:
                Kt[ny + np.arange(m), ny + np.arange(m)] -= dc
            ldu, ipiv, info = lapack.dsytrf(Kt, lower=1)
            pos, neg, zero = _inertia(ldu, ipiv) if info >= 0 else (0, 0, 1)
            if info == 0 and pos == ny and neg == m and zero == 0:
                sol, info2 = lapack.dsytrs(ldu, ipiv, rhs, lower=1)
                if info2 == 0 and np.all(np.isfinite(sol)):
                    break
                sol = None
            if zero > 0 or info > 0:
                dc = 1e-8 * mu ** 0.25
            if dw == 0.0:
                dw = 1e-4 if delta_w_last == 0.0 else max(1e-20, delta_w_last / 3.0)
            else:
                dw *= 100.0 if delta_w_last == 0.0 else 8.0
            if dw > 1e40:
                break
        if sol is None:
            break
        delta_w_last = dw
        dy, dlam = sol[:ny], sol[ny:]
        dzL = np.where(hasL, mu / dL - zL - zL / dL * dy, 0.0)
        dzU = np.where(hasU, mu / dU - zU + zU / dU * dy, 0.0)
        # ---- fraction to the boundary
        def max_step(v, dv, mask):
            neg_ = mask & (dv < 0)
            return min(1.0, float(np.min(-tau * v[neg_] / dv[neg_], initial=1.0)))

        a_pr = min(max_step(y - L, dy, hasL), max_step(U - y, -dy, hasU))
        a_du = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))
        # ---- filter line search on the barrier problem (Waechter & Biegler 2006, section 2.3), with one second-order
        #      correction when the full step is rejected because the constraint violation grew
        phi0 = barrier(y, fv, mu)
        dphi = float(gy @ dy - np.where(hasL, mu / dL, 0.0) @ dy + np.where(hasU, mu / dU, 0.0) @ dy)
        th0 = np.abs(c).sum()
        if mu != filt_mu:
            filt, filt_mu = [], mu
            th_max = 1e4 * max(1.0, th0)
            th_min = 1e-4 * max(1.0, th0)
        g_th = g_ph = 1e-5

        def acceptable(th_t, ph_t, alpha):
            if not (np.isfinite(th_t) and np.isfinite(ph_t)) or th_t > th_max:
                return False, False
            for (tf_, pf_) in filt:
                if th_t >= tf_ and ph_t >= pf_:
                    return False, False
            ftype = dphi < 0 and th0 <= th_min and alpha * (-dphi) ** 2.3 > th0 ** 1.1
            if ftype:
                return ph_t <= phi0 + 1e-4 * alpha * dphi + 10 * np.finfo(float).eps * abs(phi0), True
            return (th_t <= (1 - g_th) * th0) or (ph_t <= phi0 - g_ph * th0), False

        alpha, accepted, ftype = a_pr, False, False
        step = dy
        for ls in range(40):
            yt = y + alpha * step
            xt, st_, ft, gt, ct = evaluate(yt)
            th_t = np.abs(ct).sum() if np.all(np.isfinite(ct)) else np.inf
            ph_t = barrier(yt, ft, mu) if np.isfinite(ft) else np.inf
            accepted, ftype = acceptable(th_t, ph_t, alpha)
            if accepted:
                break
            if ls == 0 and th_t >= th0 and np.isfinite(th_t):  # second-order correction
                c_soc = alpha * c + ct
                for _ in range(4):
                    sol2, info2 = lapack.dsytrs(ldu, ipiv, -np.concatenate([r1, c_soc]), lower=1)
                    if info2 != 0 or not np.all(np.isfinite(sol2)):
                        break
                    d2 = sol2[:ny]
                    a2 = min(max_step(y - L, d2, hasL), max_step(U - y, -d2, hasU))
                    y2 = y + a2 * d2
                    x2, s2, f2, g2, c2 = evaluate(y2)
                    th2 = np.abs(c2).sum() if np.all(np.isfinite(c2)) else np.inf
                    ph2 = barrier(y2, f2, mu) if np.isfinite(f2) else np.inf
                    ok2, ft2 = acceptable(th2, ph2, a2)
                    if ok2:
                        yt, xt, st_, ft, gt, ct, alpha, step, accepted, ftype = y2, x2, s2, f2, g2, c2, a2, d2, True, ft2
                        dlam = sol2[ny:]
                        break
                    if th2 > 0.99 * th_t:
                        break
                    c_soc, th_t = a2 * c_soc + c2, th2
                if accepted:
                    break
            alpha *= 0.5
        if not accepted:  # no acceptable step at rounding level: stop here (no restoration phase in this small solver)
            break
        if not ftype:
            filt.append(((1 - g_th) * th0, phi0 - g_ph * th0))
        y, xx, s, fv, gv, c = yt, xt, st_, ft, gt, ct
        lam = lam + alpha * dlam
        zL = zL + a_du * dzL
        zU = zU + a_du * dzU
        ks = 1e10  # keep the duals within a factor of the primal estimate mu / slack
        dL = np.where(hasL, y - L, 1.0)
        dU = np.where(hasU, U - y, 1.0)
        zL = np.where(hasL, np.clip(zL, mu / (ks * dL), ks * mu / dL), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (ks * dU), ks * mu / dU), 0.0)
    # multipliers in IPOPT's convention: grad f + J^T lam_g + lam_x = 0, lam > 0 on an active upper bound
    lam_x = np.zeros(n_all)
    lam_x[free] = (zU - zL)[:n]
    if fixed.any():
        r = np.asarray(grad_f(xx), dtype=float) + sp.csr_matrix(jac_g(xx)).T @ lam
        lam_x[fixed] = -r[fixed]
    return IpmResult(x=xx, f=fv, g=gv, lam_g=lam.copy(), lam_x=lam_x, success=bool(ok), iter=it, err=float(err0), mu=mu)
