"""NLP solver front-ends that CONSUME the GPU callbacks (they are not part of the accelerated path).

The reference hands a symbolic NLP to ``ca.nlpsol(name, "ipopt", ...)`` (/root/reference/mpopt/mpopt.py:757) and
calls the returned object as ``solver(x0=, p=, lbx=, ubx=, lbg=, ubg=, lam_x0=, lam_g0=)`` getting the dict
``{x, f, g, lam_x, lam_g, lam_p}`` back (:804).  CasADi and IPOPT are not installed in this image, so:

* ``ScipyNlpSolver`` -- same call signature and result keys, driving ``Transcription.f / grad_f / g / jac_g /
  hess_l``.  Methods (``options["method"]``): ``"ipm"`` (default when the exact Hessian kernel is available and the
  KKT system is small enough to factorise densely) -- the interior-point method of ``mpopt_b200.ipm``, which reproduces
  the optimal objectives stored in the reference's notebooks to 1e-7 .. 1e-4 (tests/anchors.py); ``"SLSQP"`` and
  ``"trust-constr"`` -- SciPy's solvers (the adaptive NLP, which has no Hessian kernel, uses these).
* ``casadi_callbacks`` -- when ``import casadi`` succeeds, wraps the evaluators as ``ca.Callback`` objects with the
  Jacobian sparsity declared, ready for ``ca.nlpsol`` (untested here: no CasADi in the image).

Function evaluation is ~4 % of the reference's solve time: the solver loop itself is unchanged
by this package and stays on the host.
"""
from __future__ import annotations

import numpy as np
import scipy.optimize as so
import scipy.sparse as sp


class ScipyNlpSolver:
    def __init__(self, transcription, options=None):
        self.tr = transcription
        self.options = dict(options or {})
        self.stats = {}

    def __call__(self, x0=None, p=None, lbx=None, ubx=None, lbg=None, ubg=None, lam_x0=None, lam_g0=None):
        tr = self.tr
        x0 = np.asarray(x0, dtype=float).reshape(-1)
        p = tr.seg_width_params() if p is None else np.asarray(p, dtype=float).reshape(-1)
        lbx, ubx, lbg, ubg = (np.asarray(v, dtype=float).reshape(-1) for v in (lbx, ubx, lbg, ubg))
        rp, ci = tr.structure()
        n_z, n_g = tr.n_z, tr.n_g
        cache = {}

        def gj(z):  # one fused g + jac_g evaluation per distinct z (IPOPT's new_x contract)
            key = z.tobytes()
            if cache.get("key") != key:
                g = np.empty(n_g)
                vals = tr.jac_g_values(z, p, g_out=g)
                cache.update(key=key, g=g, J=sp.csr_matrix((vals, ci, rp), shape=(n_g, n_z)))
            return cache["g"], cache["J"]

        fobj = lambda z: tr.f(z, p)
        fgrad = lambda z: tr.grad_f(z, p)
        eq = lbg == ubg
        has_hess = True
        if getattr(tr, "adaptive", False):  # the adaptive NLP has a Hessian kernel unless something depends on t
            try:
                tr.hess_structure()
            except Exception:
                has_hess = False
        n_ineq = int((~eq).sum())
        default = "ipm" if (has_hess and n_z + n_g + n_ineq <= 6000) else ("SLSQP" if n_z <= 600 else "trust-constr")
        method = self.options.get("method", default)
        max_iter = int(self.options.get("ipopt.max_iter", self.options.get("max_iter", 500)))
        x0 = np.clip(x0, lbx, ubx)
        if method == "ipm":
            from .ipm import solve_nlp

            r = solve_nlp(fobj, fgrad, lambda z: gj(z)[0], lambda z: gj(z)[1], lambda z, lf, lg: tr.hess_l(z, p, lf, lg),
                          x0, lbx, ubx, lbg, ubg, tol=float(self.options.get("tol", self.options.get("ipopt.tol", 1e-8))),
                          max_iter=max_iter, lam_g0=lam_g0,
                          acceptable_tol=float(self.options.get("ipopt.acceptable_tol", self.options.get("acceptable_tol", 1e-4))))
            self.stats = {"success": bool(r.success), "status": "converged" if r.success else "not converged",
                          "iter_count": int(r.iter), "method": method}
            return {"x": r.x, "f": float(r.f), "g": np.asarray(r.g), "lam_x": r.lam_x, "lam_g": r.lam_g,
                    "lam_p": np.zeros(tr.n_p)}
        if method == "SLSQP":
            cons = []
            if eq.any():
                cons.append({"type": "eq", "fun": lambda z: gj(z)[0][eq] - lbg[eq],
                             "jac": lambda z: gj(z)[1][eq].toarray()})
            lo, hi = (~eq) & np.isfinite(lbg), (~eq) & np.isfinite(ubg)
            if lo.any():
                cons.append({"type": "ineq", "fun": lambda z: gj(z)[0][lo] - lbg[lo], "jac": lambda z: gj(z)[1][lo].toarray()})
            if hi.any():
                cons.append({"type": "ineq", "fun": lambda z: ubg[hi] - gj(z)[0][hi], "jac": lambda z: -gj(z)[1][hi].toarray()})
            bounds = [(l if np.isfinite(l) else None, u if np.isfinite(u) else None) for l, u in zip(lbx, ubx)]
            res = so.minimize(fobj, x0, jac=fgrad, bounds=bounds, constraints=cons, method="SLSQP",
                              options={"maxiter": max_iter, "ftol": float(self.options.get("tol", 1e-10))})
            lam_g = np.zeros(n_g)
        else:
            # exact second derivatives from the Hessian kernel (CasADi's nlp_hess_l) unless the caller asks for a
            # quasi-Newton model with IPOPT's option name
            exact = self.options.get("ipopt.hessian_approximation", self.options.get("hessian_approximation", "exact")) == "exact"
            exact = exact and has_hess
            if exact:
                zero_lam = np.zeros(n_g)

                def full(Hl):  # lower triangle -> symmetric
                    return (Hl + sp.tril(Hl, k=-1).T).tocsr()

                h_obj = lambda z: full(tr.hess_l(z, p, 1.0, zero_lam))
                h_con = lambda z, v: full(tr.hess_l(z, p, 0.0, v))
            else:
                h_obj, h_con = so.BFGS(), so.BFGS()
            nlc = so.NonlinearConstraint(lambda z: gj(z)[0], lbg, ubg, jac=lambda z: gj(z)[1], hess=h_con)
            res = so.minimize(fobj, x0, jac=fgrad, hess=h_obj, bounds=so.Bounds(lbx, ubx, keep_feasible=False),
                              constraints=[nlc], method="trust-constr",
                              options={"maxiter": max_iter, "gtol": float(self.options.get("tol", 1e-8)),
                                       "xtol": 1e-12, "sparse_jacobian": True, "verbose": 0})
            lam_g = -np.asarray(res.v[0]) if getattr(res, "v", None) else np.zeros(n_g)
        x = np.asarray(res.x, dtype=float)
        g = tr.g(x, p)
        self.stats = {"success": bool(res.success), "status": getattr(res, "message", ""), "iter_count": int(getattr(res, "nit", 0)),
                      "method": method}
        return {"x": x, "f": float(res.fun), "g": g, "lam_x": np.zeros(n_z), "lam_g": lam_g, "lam_p": np.zeros(tr.n_p)}


def casadi_callbacks(transcription):
    """(f_cb, g_cb) ``ca.Callback`` objects whose Jacobians are the GPU evaluators, for ``ca.nlpsol``.

    Only usable where CasADi is installed; kept small on purpose (the C-level route is the external-function ABI
    described in INTEGRATION.md)."""
    import casadi as ca  # noqa: F401  (ImportError here is the honest answer when CasADi is absent)

    tr = transcription
    cp, ri, perm = tr.structure_ccs()
    spJ = ca.Sparsity(tr.n_g, tr.n_z, cp.tolist(), ri.tolist())

    class JacG(ca.Callback):
        def __init__(self):
            ca.Callback.__init__(self)
            self.construct("jac_nlp_g", {})

        def get_n_in(self): return 3
        def get_n_out(self): return 2
        def get_sparsity_in(self, i): return [ca.Sparsity.dense(tr.n_z), ca.Sparsity.dense(tr.n_p), ca.Sparsity(tr.n_g, 1)][i]
        def get_sparsity_out(self, i): return [spJ, ca.Sparsity(tr.n_g, tr.n_p)][i]

        def eval(self, arg):
            vals = tr.jac_g_values(np.asarray(arg[0]).ravel(), np.asarray(arg[1]).ravel())
            return [ca.DM(spJ, vals[perm]), ca.DM(tr.n_g, tr.n_p)]

    class G(ca.Callback):
        def __init__(self):
            ca.Callback.__init__(self)
            self._jac = JacG()
            self.construct("nlp_g", {})

        def get_n_in(self): return 2
        def get_n_out(self): return 1
        def get_sparsity_in(self, i): return [ca.Sparsity.dense(tr.n_z), ca.Sparsity.dense(tr.n_p)][i]
        def get_sparsity_out(self, i): return ca.Sparsity.dense(tr.n_g)
        def eval(self, arg): return [tr.g(np.asarray(arg[0]).ravel(), np.asarray(arg[1]).ravel())]
        def has_jacobian(self): return True
        def get_jacobian(self, name, inames, onames, opts): return self._jac

    class F(ca.Callback):
        def __init__(self):
            ca.Callback.__init__(self)
            self.construct("nlp_f", {"enable_fd": False})

        def get_n_in(self): return 2
        def get_n_out(self): return 1
        def get_sparsity_in(self, i): return [ca.Sparsity.dense(tr.n_z), ca.Sparsity.dense(tr.n_p)][i]
        def get_sparsity_out(self, i): return ca.Sparsity.dense(1)
        def eval(self, arg): return [tr.f(np.asarray(arg[0]).ravel(), np.asarray(arg[1]).ravel())]

    return F(), G()
