"""Oracle (TEST INFRASTRUCTURE): the transcribed NLP  f / g / grad_f / jac_g.

CPU restatement (numpy + scipy.sparse, float64) of the reference's transcription
``mpopt`` class, /root/reference/mpopt/mpopt.py:95-639 and :641-723, plus the four
evaluators CasADi derives from it at :757 (``nlp_f``, ``nlp_g``, ``nlp_grad_f``,
``nlp_jac_g``).  Each method cites the reference lines it follows.  The Jacobian
is a ``scipy.sparse.csr_matrix`` with sorted column indices whose *pattern* is the
structural one CasADi would produce (oracle/dual.py), explicit zeros kept.

PARITY UNPINNED for g / jac_g values (no golden in the reference);
pinned for sizes (IPOPT banners), the Appendix-A golden G0, and by finite
differences (tests/test_oracle_nlp.py).
"""
from __future__ import annotations

import copy

import numpy as np
import scipy.sparse as sp

from .collocation import Tables
from .dual import Dual, Vec, flatten


def _const_or_dual(v, n):
    """-> (value array (n,), der dict, structurally_nonzero)."""
    if isinstance(v, Dual):
        val = np.broadcast_to(v.val, (n,)).astype(float)
        return val, {k: np.broadcast_to(d, (n,)).astype(float) for k, d in v.der.items()}, True
    c = float(np.asarray(v, dtype=float).reshape(-1)[0]) if not isinstance(v, (int, float)) else float(v)
    return np.full(n, c), {}, c != 0.0


class OracleNLP:
    def __init__(self, ocp, n_segments=1, poly_orders=9, scheme="LGR", tau_min=-1.0, tau_max=1.0,
                 drop_exact_zeros=True, tables=None):
        # mpopt.py:56-93
        self.K = int(n_segments)
        self.po = [poly_orders] * self.K if isinstance(poly_orders, (int, np.integer)) else list(poly_orders)
        assert len(self.po) == self.K  # :83
        self.ocp = copy.deepcopy(ocp)  # :77 (Q9)
        self.scheme = scheme
        self.N = sum(self.po) + 1  # :84
        self.tab = Tables(self.po, scheme, tau_min, tau_max, override=tables)  # :95-103
        self.tau0, self.tau1 = self.tab.tau0, self.tab.tau1
        self.drop = bool(drop_exact_zeros)
        o = self.ocp
        self.nx, self.nu, self.na, self.P = o.nx, o.nu, o.na, o.n_phases
        self.nvar = self.N * (self.nx + self.nu) + 2 + self.na  # per phase, :537-543
        self.n_z = self.P * self.nvar
        self.n_p = self.K * self.P  # :152, :631
        # node ownership, :189-195 -- a shared boundary node is the LAST point of the earlier segment
        self.seg_start = np.concatenate([[0], np.cumsum(self.po)[:-1]]).astype(np.int64)
        po = np.asarray(self.po, dtype=np.int64)
        seg = np.concatenate([[0], np.repeat(np.arange(self.K), po)]).astype(np.int64)
        loc = np.arange(self.N) - self.seg_start[seg]
        self.node_seg, self.node_loc = seg, loc
        # tau of every node inside its owning segment, minus tau0 (:198)
        self.node_dtau = np.zeros(self.N)
        for d in set(self.po):
            m = po[seg] == d
            self.node_dtau[m] = self.tab.roots[d][loc[m]] - self.tau0
        # constant composite matrices
        self._compD = self._maybe_drop(self.tab.composite_D())  # :99
        self._compW = self.tab.composite_W()  # :100
        self._compI = self._maybe_drop(self.tab.composite_mid_interpolation())  # :353-359
        self._compS = self._maybe_drop(self.tab.composite_slope_continuity()) if self.K > 1 else None  # :398-403
        # probe row structure (which optional blocks exist, how many rows each)
        self._rows = [self._phase_rows(ph) for ph in range(self.P)]
        self.n_g_phase = [r["n"] for r in self._rows]
        self.row_off = np.concatenate([[0], np.cumsum(self.n_g_phase)]).astype(np.int64)
        self.n_links = len(o.phase_links) if self.P > 1 else 0
        self.n_g = int(self.row_off[-1]) + self.n_links * (self.nx + self.nu + 1)  # :464-521, :617-621

    # ------------------------------------------------------------------ helpers
    def _maybe_drop(self, A):
        A = A.tocsr()
        if self.drop:
            A.eliminate_zeros()  # SX folds 0*x -> 0 (Q10)
        A.sort_indices()
        return A

    def colX(self, ph, i, s):
        return ph * self.nvar + s * self.N + i

    def colU(self, ph, i, c):
        return ph * self.nvar + self.nx * self.N + c * self.N + i

    def colT0(self, ph):
        return ph * self.nvar + (self.nx + self.nu) * self.N

    def colTF(self, ph):
        return self.colT0(ph) + 1

    def colA(self, ph, m):
        return self.colT0(ph) + 2 + m

    def _phase_rows(self, ph):
        o, N, K = self.ocp, self.N, self.K
        r = {}
        r["F"] = 0
        n = self.nx * N
        # :171, OCP.has_path_constraints :3573-3594
        nc = 0
        if o.has_path_constraints(ph):
            args = (o.x00[ph], o.u00[ph], o.t00[ph]) + ((o.a0[ph],) if o.na else ())
            nc = len(flatten(o.path_constraints[ph](*args)))
        r["nc"], r["C"] = nc, n
        n += nc * N
        r["DU"] = n
        r["has_DU"] = bool(o.diff_u[ph])  # :315
        n += self.nu * N if r["has_DU"] else 0
        r["mU"] = n
        # :346, :363-365
        r["has_mU"] = bool(o.midu[ph]) and bool((o.lbu[ph] > -np.inf).any() or (o.ubu[ph] < np.inf).any())
        n += self.nu * (N - 1) if r["has_mU"] else 0
        r["dU"] = n
        r["has_dU"] = (K > 1) and bool(o.du_continuity[ph])  # :394
        n += self.nu * (K - 1) if r["has_dU"] else 0
        r["TC"] = n
        ntc = 0
        if o.has_terminal_constraints(ph):  # :284, :3596-3626
            args = (o.xf0[ph], o.tf0[ph], o.x00[ph], o.t00[ph]) + ((o.a0[ph],) if o.na else ())
            ntc = len(flatten(o.terminal_constraints[ph](*args)))
        r["ntc"] = ntc
        n += ntc
        r["n"] = n
        return r

    def seg_width_params(self):
        """mpopt.py:710-723."""
        return np.array([1.0 / self.K] * (self.K * self.P))

    # ------------------------------------------------------------------ per-phase evaluation
    def _unpack(self, ph, z):
        N, nx, nu = self.N, self.nx, self.nu
        o0 = ph * self.nvar
        X = z[o0: o0 + nx * N].reshape(nx, N).T  # column-major flatten (:538) => state-major
        U = z[o0 + nx * N: o0 + (nx + nu) * N].reshape(nu, N).T
        T0, TF = z[o0 + (nx + nu) * N], z[o0 + (nx + nu) * N + 1]
        A = z[o0 + (nx + nu) * N + 2: o0 + self.nvar]
        return X, U, T0, TF, A

    def _widths(self, ph, z, p):
        """(segment widths of the phase, their columns in z or None).  Base class: widths are the NLP parameters
        ``p`` (mpopt.py:152, :631); oracle/adaptive.py makes them decision variables."""
        return np.asarray(p, dtype=float)[ph * self.K: (ph + 1) * self.K], None

    def _emit_width_cols(self, emit, grad_or_none, rows, seg, frac, coef_h, coef_t, wcols, T):
        """d/dw entries of rows whose value is  c(h_k, t)  with h_k = T/delta * w_k and t = t0 + T * sigma,
        sigma = sum_{m<k} w_m + w_k * frac:  coef_h = dc/dh * T/delta (or None), coef_t = dc/dt (or None).
        ``frac < 0`` marks a point whose time folded to t0 (node 0, mpopt.py:198): no width dependence through t."""
        if wcols is None:
            return
        seg = np.asarray(seg)
        own = np.zeros(len(seg)) if coef_h is None else np.asarray(coef_h, float).copy()
        has_t = coef_t is not None
        if has_t:
            live = frac >= 0
            own = own + np.where(live, np.asarray(coef_t, float) * T * np.where(live, frac, 0.0), 0.0)
        sel = np.ones(len(seg), bool) if coef_h is not None else (frac >= 0)
        if grad_or_none is None:
            emit(np.asarray(rows)[sel], wcols[seg[sel]], own[sel])
        else:
            np.add.at(grad_or_none, wcols[seg[sel]], own[sel])
        if has_t:  # every earlier segment shifts the point's time by T per unit width
            for i in np.nonzero(seg > 0)[0]:
                m = np.arange(seg[i])
                v = np.full(len(m), float(np.asarray(coef_t)[i]) * T)
                if grad_or_none is None:
                    emit(np.full(len(m), np.asarray(rows)[i]), wcols[m], v)
                else:
                    np.add.at(grad_or_none, wcols[m], v)

    def _time_grid(self, ph, T0, TF, w):
        """h per node, t per node and d t/d tf (sigma) -- mpopt.py:175-198."""
        o = self.ocp
        st = o.scale_t
        t0, tf = T0 / st, TF / st  # :175-176
        w = np.asarray(w, dtype=float)
        delta = self.tau1 - self.tau0
        h_seg = (tf - t0) / delta * w  # :184, :193-195
        # t_seg0 += h_seg*(tau1 - tau0), accumulated sequentially (:192)
        t_seg0 = np.empty(self.K)
        acc = t0
        for k in range(self.K):
            t_seg0[k] = acc
            acc = acc + h_seg[k] * delta
        h = h_seg[self.node_seg]
        t = t_seg0[self.node_seg] + h * self.node_dtau  # :198
        wcum = np.concatenate([[0.0], np.cumsum(w)[:-1]])
        sigma = wcum[self.node_seg] + w[self.node_seg] * self.node_dtau / delta
        dh_dtf = w[self.node_seg] / (delta * st)  # d h / d TF (scaled variable); d h / d T0 = -that
        return h, t, sigma, dh_dtf

    def _node_inputs(self, ph, X, U, A, t):
        o, N = self.ocp, self.N
        x = Vec(Dual(X[:, s] * (1.0 / o.scale_x[s]), {("x", s): np.full(N, 1.0 / o.scale_x[s])}) for s in range(self.nx))
        u = Vec(Dual(U[:, c] * (1.0 / o.scale_u[c]), {("u", c): np.full(N, 1.0 / o.scale_u[c])}) for c in range(self.nu))
        a = Vec(Dual(np.full(N, A[m] * (1.0 / o.scale_a[m])), {("a", m): np.full(N, 1.0 / o.scale_a[m])})
                for m in range(self.na))
        td = Dual(t, {("t",): np.ones(N)})
        return x, u, td, a

    def _col_of(self, ph, key, nodes):
        if key[0] == "x":
            return self.colX(ph, nodes, key[1])
        if key[0] == "u":
            return self.colU(ph, nodes, key[1])
        if key[0] == "a":
            return np.full(len(nodes), self.colA(ph, key[1]))
        raise KeyError(key)

    def _eval_phase(self, ph, z, p, want_jac=True):
        """Returns (g_phase, J_phase, (rows, cols, vals) of jac_g rows of this phase, grad_f contribution)."""
        o, N, K, nx, nu, na = self.ocp, self.N, self.K, self.nx, self.nu, self.na
        R = self._rows[ph]
        st = o.scale_t
        X, U, T0, TF, A = self._unpack(ph, z)
        w, wcols = self._widths(ph, z, p)
        h, t, sigma, dh = self._time_grid(ph, T0, TF, w)
        x, u, td, a = self._node_inputs(ph, X, U, A, t)
        nodes = np.arange(N)
        T = (TF - T0) / st
        hw = T / (self.tau1 - self.tau0)                      # d h_k / d w_k
        frac = self.node_dtau / (self.tau1 - self.tau0)       # d sigma_i / d w_k(i)
        frac_t = np.where(nodes == 0, -1.0, frac)             # node 0: t folded to t0 (:198)
        g = np.zeros(R["n"])
        rows, cols, vals = [], [], []
        grad = np.zeros(self.n_z)
        cT0, cTF = self.colT0(ph), self.colTF(ph)

        def emit(r, c, v):
            rows.append(np.asarray(r, dtype=np.int64).ravel())
            cols.append(np.asarray(c, dtype=np.int64).ravel())
            vals.append(np.asarray(v, dtype=float).ravel())

        # ---------------- dynamics defects  F = kron(I, compD) X[:] - vec(h * Sx f)   (:201, :227-232)
        fout = flatten(o.get_dynamics(ph)(x, u, td, a))
        assert len(fout) == nx
        D = self._compD
        Dcoo = D.tocoo()
        for s in range(nx):
            fv, fder, nz = _const_or_dual(fout[s], N)
            sx = o.scale_x[s]
            g[R["F"] + s * N: R["F"] + (s + 1) * N] = D @ X[:, s] - h * sx * fv
            if not want_jac:
                continue
            emit(R["F"] + s * N + Dcoo.row, self.colX(ph, Dcoo.col, s), Dcoo.data)
            ft = np.zeros(N)
            for key, d in fder.items():
                if key == ("t",):
                    ft = d
                    continue
                emit(R["F"] + s * N + nodes, self._col_of(ph, key, nodes), -h * sx * d)
            if nz:
                emit(R["F"] + s * N + nodes, np.full(N, cTF), -dh * sx * fv - h * sx * ft * sigma / st)
                emit(R["F"] + s * N + nodes, np.full(N, cT0), +dh * sx * fv - h * sx * ft * (1.0 - sigma) / st)
                self._emit_width_cols(emit, None, R["F"] + s * N + nodes, self.node_seg, frac_t, -hw * sx * fv,
                                      (-h * sx * ft) if ("t",) in fder else None, wcols, T)

        # ---------------- path constraints  C = vec(c)   (:204, :254-258)
        if R["nc"]:
            cout = flatten(o.get_path_constraints(ph)(x, u, td, a))
            assert len(cout) == R["nc"]
            for q in range(R["nc"]):
                cv, cder, _ = _const_or_dual(cout[q], N)
                r0 = R["C"] + q * N
                g[r0: r0 + N] = cv
                if not want_jac:
                    continue
                for key, d in cder.items():
                    if key == ("t",):
                        emit(r0 + nodes, np.full(N, cT0), d * (1.0 - sigma) / st)
                        # node 0: t = t0 + h*0.0 folds to t0 -> no TF dependence (:198)
                        emit(r0 + nodes[1:], np.full(N - 1, cTF), (d * sigma / st)[1:])
                        self._emit_width_cols(emit, None, r0 + nodes, self.node_seg, frac_t, None, d, wcols, T)
                    else:
                        emit(r0 + nodes, self._col_of(ph, key, nodes), d)

        # ---------------- control slope  kron(I, compD) U[:]   (:315-324)
        if R["has_DU"]:
            for c in range(nu):
                r0 = R["DU"] + c * N
                g[r0: r0 + N] = D @ U[:, c]
                if want_jac:
                    emit(r0 + Dcoo.row, self.colU(ph, Dcoo.col, c), Dcoo.data)

        # ---------------- mid-point control box  (:346-375)
        if R["has_mU"]:
            Icoo = self._compI.tocoo()
            for c in range(nu):
                r0 = R["mU"] + c * (N - 1)
                g[r0: r0 + N - 1] = self._compI @ U[:, c]
                if want_jac:
                    emit(r0 + Icoo.row, self.colU(ph, Icoo.col, c), Icoo.data)

        # ---------------- slope continuity across segments  (:394-411)
        if R["has_dU"]:
            Scoo = self._compS.tocoo()
            for c in range(nu):
                r0 = R["dU"] + c * (K - 1)
                g[r0: r0 + K - 1] = self._compS @ U[:, c]
                if want_jac:
                    emit(r0 + Scoo.row, self.colU(ph, Scoo.col, c), Scoo.data)

        # ---------------- terminal constraints and Mayer term  (:277-298)
        one = np.ones(1)
        x0 = Vec(Dual(X[0:1, s] * (1.0 / o.scale_x[s]), {("x0", s): one / o.scale_x[s]}) for s in range(nx))
        xf = Vec(Dual(X[N - 1: N, s] * (1.0 / o.scale_x[s]), {("xf", s): one / o.scale_x[s]}) for s in range(nx))
        a1 = Vec(Dual(A[m: m + 1] * (1.0 / o.scale_a[m]), {("a", m): one / o.scale_a[m]}) for m in range(na))
        t0d = Dual(np.array([T0 / st]), {("t0",): one / st})
        tfd = Dual(np.array([TF / st]), {("tf",): one / st})

        def tcol(key):
            return {"x0": lambda: self.colX(ph, 0, key[1]), "xf": lambda: self.colX(ph, N - 1, key[1]),
                    "a": lambda: self.colA(ph, key[1]), "t0": lambda: cT0, "tf": lambda: cTF}[key[0]]()

        if R["ntc"]:
            tc = flatten(o.get_terminal_constraints(ph)(xf, tfd, x0, t0d, a1))
            assert len(tc) == R["ntc"]
            for r_, e in enumerate(tc):
                v, der, _ = _const_or_dual(e, 1)
                g[R["TC"] + r_] = v[0]
                if want_jac:
                    for key, d in der.items():
                        emit([R["TC"] + r_], [tcol(key)], d)
        M = flatten(o.get_terminal_costs(ph)(xf, tfd, x0, t0d, a1))
        Mv, Mder, _ = _const_or_dual(M[0], 1)
        J = Mv[0]
        for key, d in Mder.items():
            grad[tcol(key)] += d[0]

        # ---------------- running cost  J += compW . (h L)   (:206, :455)
        Lout = flatten(o.get_running_costs(ph)(x, u, td, a))
        Lv, Lder, Lnz = _const_or_dual(Lout[0], N)
        W = self._compW
        J = J + float(W @ (h * Lv))
        Lt = np.zeros(N)
        for key, d in Lder.items():
            if key == ("t",):
                Lt = d
            elif key[0] == "a":
                grad[self.colA(ph, key[1])] += float(W @ (h * d))
            else:
                grad[self._col_of(ph, key, nodes)] += W * h * d
        if Lnz:
            grad[cTF] += float(W @ (dh * Lv + h * Lt * sigma / st))
            grad[cT0] += float(W @ (-dh * Lv + h * Lt * (1.0 - sigma) / st))
            self._emit_width_cols(None, grad, nodes, self.node_seg, frac_t, W * hw * Lv,
                                  (W * h * Lt) if ("t",) in Lder else None, wcols, T)

        self._extra_rows(ph, R, g, emit if want_jac else None, X, U, T0, TF, A, w, wcols, t)
        trip = (np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)) if rows else (
            np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0))
        return g, J, trip, grad

    def _extra_rows(self, ph, R, g, emit, X, U, T0, TF, A, w, wcols, t):
        """Hook for subclasses that append constraint blocks to a phase (oracle/adaptive.py)."""

    def _events(self, z, want_jac=True):
        """Phase-link rows (mpopt.py:464-521), appended after all phases (:617-621)."""
        o, N, nx, nu = self.ocp, self.N, self.nx, self.nu
        n = self.n_links
        g = np.zeros(n * (nx + nu + 1))
        rows, cols, vals = [], [], []
        base = int(self.row_off[-1])
        for li, (pi, pj) in enumerate(o.phase_links):
            for s in range(nx):
                r = li * nx + s
                g[r] = z[self.colX(pj, 0, s)] - z[self.colX(pi, N - 1, s)]
                rows += [base + r, base + r]; cols += [self.colX(pj, 0, s), self.colX(pi, N - 1, s)]; vals += [1.0, -1.0]
            for c in range(nu):
                r = n * nx + li * nu + c
                g[r] = z[self.colU(pj, 0, c)] - z[self.colU(pi, N - 1, c)]
                rows += [base + r, base + r]; cols += [self.colU(pj, 0, c), self.colU(pi, N - 1, c)]; vals += [1.0, -1.0]
            r = n * (nx + nu) + li
            g[r] = z[self.colT0(pj)] - z[self.colTF(pi)]
            rows += [base + r, base + r]; cols += [self.colT0(pj), self.colTF(pi)]; vals += [1.0, -1.0]
        return g, (np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, float))

    # ------------------------------------------------------------------ public evaluators
    def _eval(self, z, p=None, want_jac=True):
        z = np.asarray(z, dtype=float)
        p = self.seg_width_params() if p is None else np.asarray(p, dtype=float)
        assert z.shape == (self.n_z,) and p.shape == (self.n_p,)
        gs, J, grad = [], 0.0, np.zeros(self.n_z)
        R, C, V = [], [], []
        for ph in range(self.P):
            g, Jp, (r, c, v), gr = self._eval_phase(ph, z, p, want_jac)
            gs.append(g)
            J += Jp  # :615
            grad += gr
            R.append(r + self.row_off[ph]); C.append(c); V.append(v)
        if self.n_links:
            g, (r, c, v) = self._events(z)
            gs.append(g); R.append(r); C.append(c); V.append(v)
        g = np.concatenate(gs)
        jac = None
        if want_jac:
            jac = sp.coo_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(C))),
                                shape=(self.n_g, self.n_z)).tocsr()  # sums duplicates, keeps explicit zeros
            jac.sort_indices()
        return J, g, grad, jac

    def f(self, z, p=None):
        return self._eval(z, p, want_jac=False)[0]

    def g(self, z, p=None):
        return self._eval(z, p, want_jac=False)[1]

    def grad_f(self, z, p=None):
        return self._eval(z, p, want_jac=False)[2]

    def jac_g(self, z, p=None):
        return self._eval(z, p)[3]

    def structure(self, z=None):
        """(rowptr, colind) of jac_g as int64, sorted columns; pattern does not depend on z."""
        if z is None:
            z = np.random.default_rng(0).uniform(0.5, 1.5, self.n_z)
        J = self.jac_g(z)
        return J.indptr.astype(np.int64), J.indices.astype(np.int64)

    # ------------------------------------------------------------------ bounds and initial guess
    def bounds(self):
        """(Zmin, Zmax, Gmin, Gmax) -- mpopt.py:546-570, :234-235, :257-258, :291-292, :323-324, :368-369,
        :410-411, :491-519."""
        o, N, K, nx, nu = self.ocp, self.N, self.K, self.nx, self.nu
        Zmin, Zmax, Gmin, Gmax = [], [], [], []
        for ph in range(self.P):
            xmin = [o.lbx[ph] * o.scale_x] * N
            xmax = [o.ubx[ph] * o.scale_x] * N
            if ph == 0:
                xmin[0] = xmax[0] = o.x00[0] * o.scale_x  # :550-551
            Zmin.append(np.concatenate([
                np.concatenate(np.array(xmin).T) if nx else np.zeros(0),
                np.repeat(o.lbu[ph] * o.scale_u, N),
                np.atleast_1d(o.lbt0[ph] * o.scale_t), np.atleast_1d(o.lbtf[ph] * o.scale_t),
                o.lba[ph] * o.scale_a]))
            Zmax.append(np.concatenate([
                np.concatenate(np.array(xmax).T) if nx else np.zeros(0),
                np.repeat(o.ubu[ph] * o.scale_u, N),
                np.atleast_1d(o.ubt0[ph] * o.scale_t), np.atleast_1d(o.ubtf[ph] * o.scale_t),
                o.uba[ph] * o.scale_a]))
            R = self._rows[ph]
            lo = [np.full(nx * N, float(o.LB_DYNAMICS)), np.full(R["nc"] * N, float(o.LB_PATH_CONSTRAINTS))]
            hi = [np.full(nx * N, float(o.UB_DYNAMICS)), np.full(R["nc"] * N, float(o.UB_PATH_CONSTRAINTS))]
            if R["has_DU"]:
                lo.append(np.full(nu * N, float(o.lbdu[ph]))); hi.append(np.full(nu * N, float(o.ubdu[ph])))
            if R["has_mU"]:
                lo.append(np.repeat(o.lbu[ph] * o.scale_u, N - 1)); hi.append(np.repeat(o.ubu[ph] * o.scale_u, N - 1))
            if R["has_dU"]:
                lo.append(np.zeros(nu * (K - 1))); hi.append(np.zeros(nu * (K - 1)))
            lo.append(np.full(R["ntc"], float(o.LB_TERMINAL_CONSTRAINTS)))
            hi.append(np.full(R["ntc"], float(o.UB_TERMINAL_CONSTRAINTS)))
            Gmin.append(np.concatenate(lo)); Gmax.append(np.concatenate(hi))
        if self.n_links:
            n = self.n_links
            Gmin.append(np.concatenate([o.lbe[i] * o.scale_x for i in range(n)]))  # Q5: indexed by link ordinal
            Gmax.append(np.concatenate([o.ube[i] * o.scale_x for i in range(n)]))
            Gmin.append(np.zeros(n * nu)); Gmax.append(np.zeros(n * nu))
            Gmin.append(np.zeros(n)); Gmax.append(np.zeros(n))
        return (np.concatenate(Zmin).astype(float), np.concatenate(Zmax).astype(float),
                np.concatenate(Gmin).astype(float), np.concatenate(Gmax).astype(float))

    def initialize_solution(self):
        """mpopt.py:641-708 (X state-major, U node-major -- quirk Q3)."""
        o, N = self.ocp, self.N
        Z0 = []
        for ph in range(self.P):
            x00, xf0 = o.x00[ph] * o.scale_x, o.xf0[ph] * o.scale_x
            u00, uf0 = o.u00[ph] * o.scale_u, o.uf0[ph] * o.scale_u
            t00, tf0 = o.t00[ph] * o.scale_t, o.tf0[ph] * o.scale_t
            a0 = o.a0[ph] * o.scale_a
            ts = np.linspace(t00, tf0, N)
            zx = np.concatenate(np.array([x00 + (xf0 - x00) / (tf0 - t00) * (t - t00) for t in ts]).T)
            zu = np.concatenate(np.array([u00 + (uf0 - u00) / (tf0 - t00) * (t - t00) for t in ts]))
            Z0.append(np.concatenate([zx, zu, np.atleast_1d(t00), np.atleast_1d(tf0), a0]))
        return np.concatenate(Z0).astype(float)
