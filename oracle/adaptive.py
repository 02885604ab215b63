"""Oracle (TEST INFRASTRUCTURE): the NLP of ``mpopt_adaptive`` -- segment widths as decision variables.

CPU restatement (numpy, float64) of /root/reference/mpopt/mpopt.py:2877-3375 on top of oracle/nlp.py:

* variables per phase ``[X, U, t0, tf, a, w_0 .. w_{K-1}]`` (:2927-2979), no NLP parameters (``p`` is popped, :3190-3191);
* rows per phase ``[F, C, DU, TC, SW]`` (:3169) -- the mid-point control rows and the slope-continuity rows of the
  base class are NOT part of this NLP;
* ``SW`` (:3034-3136) = ``[sum(w) - 1 | compI.U (if any control bound is finite) | compI.X (if any state bound is
  finite) | mid-point residuals (if mid_residuals)]`` with the residual of segment k, state s, mid point m
  ``w_k * (DI_k[m, :] . X(seg k, s) - h_k * sx_s * f_s(xi / sx, ui / su, ti, a))`` ordered segment-major, then
  state-major (``[:]`` of a ``p_k x nx`` matrix, :3117-3120), ``ti`` the average of the two neighbouring node times
  (:3048-3053), bounds ``+-tol_residual``;
* initial guess: the base guess plus equal widths (:2981-3032), width bounds ``[lbh, ubh]`` (:2958-2976).

PARITY UNPINNED for values (no fixture in the reference); the Jacobian is checked by finite differences and the sizes
by the row formulas above (tests/test_oracle_adaptive.py).
"""
from __future__ import annotations

import numpy as np

from .collocation import diff_matrix, interpolation_matrix
from .dual import Dual, Vec, flatten
from .nlp import OracleNLP, _const_or_dual


class OracleAdaptiveNLP(OracleNLP):
    _SEG_WIDTH_MIN = 1e-4  # :2896-2898
    _SEG_WIDTH_MAX = 1.0
    _TOL_RESIDUAL = 1e-3

    def __init__(self, ocp, n_segments=1, poly_orders=9, scheme="LGR", tau_min=-1.0, tau_max=1.0,
                 drop_exact_zeros=True, tables=None, mid_residuals=True):
        self.mid_residuals = bool(mid_residuals)  # :2918
        super().__init__(ocp, n_segments, poly_orders, scheme, tau_min, tau_max, drop_exact_zeros, tables)
        self.nvar += self.K  # :2945 -- widths appended after the parameters
        self.n_z = self.P * self.nvar
        self.n_p = 0  # :3190-3191
        self.lbh = [self._SEG_WIDTH_MIN] * self.P
        self.ubh = [self._SEG_WIDTH_MAX] * self.P
        self.tol_residual = [self._TOL_RESIDUAL] * self.P
        # per-segment mid-point tables (:3040-3046, :3055-3060): C at the mid points, D at the mid points
        self._mid = {}
        for d in set(self.po):
            r = self.tab.roots[d]
            mid = 0.5 * (r[:-1] + r[1:])
            self._mid[d] = (mid, interpolation_matrix(r, mid), diff_matrix(r, mid, 1))

    def colW(self, ph, k):
        return ph * self.nvar + (self.nx + self.nu) * self.N + 2 + self.na + np.asarray(k)

    def _phase_rows(self, ph):
        o, N = self.ocp, self.N
        r = super()._phase_rows(ph)
        # rebuild the offsets: [F, C, DU, TC, SW]
        n = self.nx * N + r["nc"] * N
        r["DU"] = n
        n += self.nu * N if r["has_DU"] else 0
        r["has_mU"] = r["has_dU"] = False
        r["mU"] = r["dU"] = n
        r["TC"] = n
        n += r["ntc"]
        r["SW"] = n
        r["sw_u"] = bool((o.lbu[ph] > -np.inf).any() or (o.ubu[ph] < np.inf).any())  # :3066-3068
        r["sw_x"] = bool((o.lbx[ph] > -np.inf).any() or (o.ubx[ph] < np.inf).any())  # :3075-3077
        n += 1 + (self.nu * (N - 1) if r["sw_u"] else 0) + (self.nx * (N - 1) if r["sw_x"] else 0)
        n += self.nx * (N - 1) if self.mid_residuals else 0
        r["n"] = n
        return r

    def seg_width_params(self):
        return np.zeros(0)

    def _widths(self, ph, z, p):
        cols = self.colW(ph, np.arange(self.K))
        return np.asarray(z, dtype=float)[cols], cols

    def _unpack(self, ph, z):
        X, U, T0, TF, _ = super()._unpack(ph, z)
        o0 = ph * self.nvar + (self.nx + self.nu) * self.N + 2
        return X, U, T0, TF, z[o0: o0 + self.na]

    # ------------------------------------------------------------------ the SW block (:3034-3136)
    def _extra_rows(self, ph, R, g, emit, X, U, T0, TF, A, w, wcols, t):
        o, N, K, nx, nu, na = self.ocp, self.N, self.K, self.nx, self.nu, self.na
        st = o.scale_t
        delta = self.tau1 - self.tau0
        T = (TF - T0) / st
        r0 = R["SW"]
        # ---- sum of the widths (:3038)
        g[r0] = np.sum(w) - 1.0
        if emit:
            emit(np.full(K, r0), wcols, np.ones(K))
        r0 += 1
        Icoo = self._compI.tocoo()
        # ---- controls, then states, at the mid points (:3062-3082)
        if R["sw_u"]:
            for c in range(nu):
                g[r0: r0 + N - 1] = self._compI @ U[:, c]
                if emit:
                    emit(r0 + Icoo.row, self.colU(ph, Icoo.col, c), Icoo.data)
                r0 += N - 1
        if R["sw_x"]:
            for s in range(nx):
                g[r0: r0 + N - 1] = self._compI @ X[:, s]
                if emit:
                    emit(r0 + Icoo.row, self.colX(ph, Icoo.col, s), Icoo.data)
                r0 += N - 1
        if not self.mid_residuals:
            return
        # ---- mid-point residuals (:3084-3124)
        nm = N - 1
        seg = np.repeat(np.arange(K), self.po)                # segment of every mid point
        mloc = np.arange(nm) - self.seg_start[seg]
        xi = np.empty((nm, nx)); ui = np.empty((nm, nu)); dxi = np.empty((nm, nx))
        for k, d in enumerate(self.po):
            s0 = self.seg_start[k]
            _, Cm, Dm = self._mid[d]
            xi[s0: s0 + d] = Cm @ X[s0: s0 + d + 1]
            ui[s0: s0 + d] = Cm @ U[s0: s0 + d + 1]
            dxi[s0: s0 + d] = Dm @ X[s0: s0 + d + 1]
        ti = 0.5 * (t[:-1] + t[1:])  # :3048-3053
        frac = 0.5 * (self.node_dtau[:-1] * (self.node_seg[:-1] == seg) + self.node_dtau[1:]) / delta
        wcum = np.concatenate([[0.0], np.cumsum(w)[:-1]])
        sigma = wcum[seg] + w[seg] * frac
        hk = T / delta * w[seg]
        xd = Vec(Dual(xi[:, s] * (1.0 / o.scale_x[s]), {("x", s): np.full(nm, 1.0 / o.scale_x[s])}) for s in range(nx))
        ud = Vec(Dual(ui[:, c] * (1.0 / o.scale_u[c]), {("u", c): np.full(nm, 1.0 / o.scale_u[c])}) for c in range(nu))
        ad = Vec(Dual(np.full(nm, A[m] * (1.0 / o.scale_a[m])), {("a", m): np.full(nm, 1.0 / o.scale_a[m])})
                 for m in range(na))
        fout = flatten(o.get_dynamics(ph)(xd, ud, Dual(ti, {("t",): np.ones(nm)}), ad))
        # row of (segment k, state s, local mid point m): segments first, then states (:3117-3120)
        po = np.asarray(self.po)
        row = r0 + nx * self.seg_start[seg] + mloc            # + s * p_k below
        cT0, cTF = self.colT0(ph), self.colTF(ph)
        for s in range(nx):
            fv, fder, nz = _const_or_dual(fout[s], nm)
            sx = o.scale_x[s]
            rows_s = row + s * po[seg]
            body = dxi[:, s] - hk * sx * fv
            g[rows_s] = w[seg] * body
            if not emit:
                continue
            # d/dX through DI (state s) and through xi (every state f_s depends on); d/dU through ui
            for k, d in enumerate(self.po):
                s0 = self.seg_start[k]
                _, Cm, Dm = self._mid[d]
                rk = rows_s[s0: s0 + d]
                cols = s0 + np.arange(d + 1)
                Dk = w[k] * Dm
                mask = (Dk != 0.0) if (self.drop and ("x", s) not in fder) else np.ones_like(Dk, bool)
                if ("x", s) in fder:
                    Dk = Dk - (w[k] * hk[s0: s0 + d] * sx * fder[("x", s)][s0: s0 + d])[:, None] * Cm
                rr, cc = np.nonzero(mask)
                emit(rk[rr], self.colX(ph, cols[cc], s), Dk[rr, cc])
                for key, dv in fder.items():
                    if key[0] not in ("x", "u") or key == ("x", s):
                        continue
                    blk = -(w[k] * hk[s0: s0 + d] * sx * dv[s0: s0 + d])[:, None] * Cm
                    colf = self.colX if key[0] == "x" else self.colU
                    emit(np.repeat(rk, d + 1), np.tile(colf(ph, cols, key[1]), d), blk.ravel())
            ft = fder.get(("t",), None)
            for key, dv in fder.items():
                if key[0] == "a":
                    emit(rows_s, np.full(nm, self.colA(ph, key[1])), -w[seg] * hk * sx * dv)
            ftv = np.zeros(nm) if ft is None else ft
            if nz:
                dh = w[seg] / (delta * st)
                emit(rows_s, np.full(nm, cTF), w[seg] * (-dh * sx * fv - hk * sx * ftv * sigma / st))
                emit(rows_s, np.full(nm, cT0), w[seg] * (+dh * sx * fv - hk * sx * ftv * (1.0 - sigma) / st))
            # widths: the factor w_k itself, h_k, and (time-dependent dynamics) the mid-point time
            coef_h = body + (w[seg] * (-(T / delta) * sx * fv) if nz else 0.0)
            self._emit_width_cols(emit, None, rows_s, seg, frac, coef_h,
                                  (-w[seg] * hk * sx * ft) if ft is not None else None, wcols, T)

    # ------------------------------------------------------------------ bounds and initial guess
    def bounds(self):
        """:2948-2979 (variables) and :3169-3172 with :3037-3130 (rows)."""
        o, N, K, nx, nu = self.ocp, self.N, self.K, self.nx, self.nu
        Zmin, Zmax, Gmin, Gmax = [], [], [], []
        for ph in range(self.P):
            xmin = [o.lbx[ph] * o.scale_x] * N
            xmax = [o.ubx[ph] * o.scale_x] * N
            if ph == 0:
                xmin[0] = xmax[0] = o.x00[0] * o.scale_x
            Zmin.append(np.concatenate([
                np.concatenate(np.array(xmin).T), np.repeat(o.lbu[ph] * o.scale_u, N),
                np.atleast_1d(o.lbt0[ph] * o.scale_t), np.atleast_1d(o.lbtf[ph] * o.scale_t),
                o.lba[ph] * o.scale_a, [self.lbh[ph]] * K]))
            Zmax.append(np.concatenate([
                np.concatenate(np.array(xmax).T), np.repeat(o.ubu[ph] * o.scale_u, N),
                np.atleast_1d(o.ubt0[ph] * o.scale_t), np.atleast_1d(o.ubtf[ph] * o.scale_t),
                o.uba[ph] * o.scale_a, [self.ubh[ph]] * K]))
            R = self._rows[ph]
            lo = [np.full(nx * N, float(o.LB_DYNAMICS)), np.full(R["nc"] * N, float(o.LB_PATH_CONSTRAINTS))]
            hi = [np.full(nx * N, float(o.UB_DYNAMICS)), np.full(R["nc"] * N, float(o.UB_PATH_CONSTRAINTS))]
            if R["has_DU"]:
                lo.append(np.full(nu * N, float(o.lbdu[ph]))); hi.append(np.full(nu * N, float(o.ubdu[ph])))
            lo.append(np.full(R["ntc"], float(o.LB_TERMINAL_CONSTRAINTS)))
            hi.append(np.full(R["ntc"], float(o.UB_TERMINAL_CONSTRAINTS)))
            lo.append(np.zeros(1)); hi.append(np.zeros(1))
            if R["sw_u"]:
                lo.append(np.repeat(o.lbu[ph] * o.scale_u, N - 1)); hi.append(np.repeat(o.ubu[ph] * o.scale_u, N - 1))
            if R["sw_x"]:
                lo.append(np.repeat(o.lbx[ph] * o.scale_x, N - 1)); hi.append(np.repeat(o.ubx[ph] * o.scale_x, N - 1))
            if self.mid_residuals:
                lo.append(np.full(nx * (N - 1), -self.tol_residual[ph]))
                hi.append(np.full(nx * (N - 1), self.tol_residual[ph]))
            Gmin.append(np.concatenate(lo)); Gmax.append(np.concatenate(hi))
        if self.n_links:
            n = self.n_links
            Gmin.append(np.concatenate([o.lbe[i] * o.scale_x for i in range(n)]))
            Gmax.append(np.concatenate([o.ube[i] * o.scale_x for i in range(n)]))
            Gmin.append(np.zeros(n * nu)); Gmax.append(np.zeros(n * nu))
            Gmin.append(np.zeros(n)); Gmax.append(np.zeros(n))
        return (np.concatenate(Zmin).astype(float), np.concatenate(Zmax).astype(float),
                np.concatenate(Gmin).astype(float), np.concatenate(Gmax).astype(float))

    def initialize_solution(self):
        """:2981-3032 -- the base guess per phase followed by equal widths."""
        base = OracleNLP.initialize_solution(self)
        nb = self.nvar - self.K
        return np.concatenate([np.concatenate([base[ph * nb: (ph + 1) * nb], np.full(self.K, 1.0 / self.K)])
                               for ph in range(self.P)])

    # ------------------------------------------------------------------ Hessian of the Lagrangian (oracle/hessian.py)
    def _time_dependent(self, ph):
        """True if the dynamics, path constraints or running cost of the phase depend on t explicitly."""
        o, one = self.ocp, np.ones(1)
        x = Vec(Dual(0.7 * one, {("x", s): one}) for s in range(self.nx))
        u = Vec(Dual(0.6 * one, {("u", c): one}) for c in range(self.nu))
        a = Vec(Dual(0.5 * one, {("a", m): one}) for m in range(self.na))
        t = Dual(0.4 * one, {("t",): one})
        outs = flatten(o.get_dynamics(ph)(x, u, t, a)) + flatten(o.get_running_costs(ph)(x, u, t, a))
        if self._rows[ph]["nc"]:
            outs += flatten(o.get_path_constraints(ph)(x, u, t, a))
        return any(isinstance(e, Dual) and ("t",) in e.der for e in outs)

    def _extra_hessian(self, ph, z, lam_phase, rows, cols, vals):
        """Second derivatives of  sum lam_R R  over the mid-point residual rows
        R(k, s, m) = w_k (DI_k[m,:] X(seg k, s) - h_k sx_s f_s(CI_k[m,:] X / sx, CI_k[m,:] U / su, t_m, a / sa)):
        every node variable of a segment couples with every other one of that segment, with w_k, T0, TF and a.
        (sum(w) - 1, compI.U and compI.X are linear.)  Time-dependent problems are refused: there t_m, and in the base
        rows t_i, also depend on every earlier width, which hess_l's node-local variables do not model."""
        from .dual2 import Dual2

        if self._time_dependent(ph):
            raise NotImplementedError("Hessian oracle of the adaptive NLP: explicit time dependence is not covered")
        if not self.mid_residuals:
            return
        o, N, K, nx, nu, na = self.ocp, self.N, self.K, self.nx, self.nu, self.na
        R = self._rows[ph]
        X, U, T0, TF, A = self._unpack(ph, z)
        w, wcols = self._widths(ph, z, None)
        st, delta = o.scale_t, self.tau1 - self.tau0
        r0 = R["SW"] + 1 + (nu * (N - 1) if R["sw_u"] else 0) + (nx * (N - 1) if R["sw_x"] else 0)
        for k, d in enumerate(self.po):
            s0 = int(self.seg_start[k])
            _, Cm, Dm = self._mid[d]
            one = np.ones(d)
            Xl = [[Dual2.variable(X[s0 + j, s] * one, ("x", j, s)) for s in range(nx)] for j in range(d + 1)]
            Ul = [[Dual2.variable(U[s0 + j, c] * one, ("u", j, c)) for c in range(nu)] for j in range(d + 1)]
            Ad = [Dual2.variable(A[m] * one, ("a", m)) for m in range(na)]
            T0d, TFd, Wd = Dual2.variable(T0 * one, ("T0",)), Dual2.variable(TF * one, ("TF",)), Dual2.variable(w[k] * one, ("w",))

            def interp(M, loc, q):
                acc = loc[0][q] * M[:, 0]
                for j in range(1, d + 1):
                    acc = acc + loc[j][q] * M[:, j]
                return acc

            xi = Vec(interp(Cm, Xl, s) * (1.0 / o.scale_x[s]) for s in range(nx))
            ui = Vec(interp(Cm, Ul, c) * (1.0 / o.scale_u[c]) for c in range(nu))
            a = Vec(Ad[m] * (1.0 / o.scale_a[m]) for m in range(na))
            hk = (TFd - T0d) * (1.0 / (st * delta)) * Wd
            f = o.get_dynamics(ph)(xi, ui, Dual2(np.zeros(d)), a)
            f = list(f) if isinstance(f, (list, tuple)) else [f]
            lag = Dual2(np.zeros(d))
            for s in range(nx):
                lam_s = lam_phase[r0 + nx * s0 + s * d: r0 + nx * s0 + (s + 1) * d]
                lag = lag + (Wd * (interp(Dm, Xl, s) - hk * f[s] * o.scale_x[s])) * lam_s

            def col(key):
                if key[0] == "x":
                    return self.colX(ph, s0 + key[1], key[2])
                if key[0] == "u":
                    return self.colU(ph, s0 + key[1], key[2])
                if key[0] == "a":
                    return self.colA(ph, key[1])
                return {"T0": self.colT0(ph), "TF": self.colTF(ph), "w": int(wcols[k])}[key[0]]

            for (ka, kb), v in lag.H.items():
                ca, cb = col(ka), col(kb)
                rows.append(np.array([max(ca, cb)])), cols.append(np.array([min(ca, cb)])), vals.append(np.array([np.sum(v)]))
