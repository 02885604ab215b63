"""Oracle (TEST INFRASTRUCTURE): Hessian of the Lagrangian, lower triangle, CSR with sorted columns.

CPU restatement of CasADi's ``nlp_hess_l(x, p, lam_f, lam_g)`` for the reference's transcription: the function IPOPT
calls through ``ca.nlpsol`` (implicit in /root/reference/mpopt/mpopt.py:757; it is the third-largest evaluator in every
stored timing table, e.g. docs/source/notebooks/multi_stage_launch_vehicle_ascent.ipynb:503).  The Lagrangian is
``lam_f * J + lam_g . G`` with J and G exactly as oracle/nlp.py restates them (mpopt.py:154-462); only the terms that
are non-linear in the decision vector contribute:

    running cost    lam_f * compW_i * h_k L(x_i, u_i, t_i, a)            (:206, :455)
    defect rows     - lam_F(s,i) * h_k * Sx_s * f_s(x_i, u_i, t_i, a)     (:201, :232)
    path rows       lam_C(q,i) * c_q(x_i, u_i, t_i, a)                    (:204)
    Mayer term / terminal rows   lam_f * M(.) + lam_TC(r) * tc_r(xf, tf, x0, t0, a)   (:277-298)

with h_k and t_i functions of (T0, TF) (:175-198).  Second derivatives come from oracle/dual2.py applied to the RAW
decision variables (scaling and the time map are part of the differentiated expression), so the chain rule is not
restated by hand.  A pair is in the pattern iff it is structurally non-zero.  PARITY UNPINNED: the reference holds no
value or pattern of nlp_hess_l; cross-checked by finite differences of the first-order oracle (tests/test_oracle_hessian.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .dual import Vec
from .dual2 import Dual2


class _V(Vec):
    """What the user callables index as x[i] / u[i] / a[i] (slices and scalar x vector products included)."""


def _components(out, n):
    out = list(out) if isinstance(out, (list, tuple)) else [out]
    assert len(out) == n
    return out


def hess_l(ora, z, p=None, lam_f=1.0, lam_g=None):
    """Lower triangle of the Lagrangian Hessian at z: scipy.sparse.csr_matrix (n_z, n_z), explicit structural zeros kept."""
    o, N, K, nx, nu, na = ora.ocp, ora.N, ora.K, ora.nx, ora.nu, ora.na
    z = np.asarray(z, dtype=float)
    p = ora.seg_width_params() if p is None else np.asarray(p, dtype=float)
    lam = np.zeros(ora.n_g) if lam_g is None else np.asarray(lam_g, dtype=float)
    st = o.scale_t
    rows, cols, vals = [], [], []

    for ph in range(ora.P):
        R = ora._rows[ph]
        base = int(ora.row_off[ph])
        X, U, T0, TF, A = ora._unpack(ph, z)
        w, wcols = ora._widths(ph, z, p)  # parameters, or (oracle/adaptive.py) decision variables with columns wcols
        delta = ora.tau1 - ora.tau0
        _, _, sigma, _ = ora._time_grid(ph, T0, TF, w)
        wn = w[ora.node_seg]
        if wcols is not None:  # h_k = (tf - t0)/delta * w_k is bilinear in (T0 | TF, w_k); sigma is handled by the subclass
            wn = Dual2.variable(wn, ("w",))
        ones = np.ones(N)
        # raw decision variables as second-order duals, one entry per node
        Xd = [Dual2.variable(X[:, s], ("x", s)) for s in range(nx)]
        Ud = [Dual2.variable(U[:, c], ("u", c)) for c in range(nu)]
        Ad = [Dual2.variable(A[m] * ones, ("a", m)) for m in range(na)]
        T0d, TFd = Dual2.variable(T0 * ones, ("T0",)), Dual2.variable(TF * ones, ("TF",))
        x = _V(Xd[s] * (1.0 / o.scale_x[s]) for s in range(nx))
        u = _V(Ud[c] * (1.0 / o.scale_u[c]) for c in range(nu))
        a = _V(Ad[m] * (1.0 / o.scale_a[m]) for m in range(na))
        t0, tf = T0d * (1.0 / st), TFd * (1.0 / st)  # :175-176
        h = (tf - t0) * (wn * (1.0 / delta))  # :184
        t = t0 + (tf - t0) * sigma  # :192, :198
        lag = Dual2(np.zeros(N))
        f = _components(o.get_dynamics(ph)(x, u, t, a), nx)
        for s in range(nx):
            lF = lam[base + R["F"] + s * N: base + R["F"] + (s + 1) * N]
            lag = lag - (h * f[s]) * (lF * o.scale_x[s])
        if R["nc"]:
            c = _components(o.get_path_constraints(ph)(x, u, t, a), R["nc"])
            for q in range(R["nc"]):
                lC = lam[base + R["C"] + q * N: base + R["C"] + (q + 1) * N]
                lag = lag + c[q] * lC if isinstance(c[q], Dual2) else lag
        L = o.get_running_costs(ph)(x, u, t, a)
        L = L[0] if isinstance(L, (list, tuple)) else L
        if isinstance(L, Dual2) or float(np.asarray(L).reshape(-1)[0]) != 0.0:
            lag = lag + (h * L) * (lam_f * ora._compW)
        nodes = np.arange(N)

        def ncol(key):
            if key[0] == "x":
                return ora.colX(ph, nodes, key[1])
            if key[0] == "u":
                return ora.colU(ph, nodes, key[1])
            if key[0] == "a":
                return ora.colA(ph, key[1])
            if key[0] == "w":
                return wcols[ora.node_seg]
            return ora.colT0(ph) if key[0] == "T0" else ora.colTF(ph)

        for (ka, kb), v in lag.H.items():
            ca, cb = ncol(ka), ncol(kb)
            if np.ndim(ca) == 0 and np.ndim(cb) == 0:  # both global variables: one entry, summed over the nodes
                rows.append(np.array([max(ca, cb)])), cols.append(np.array([min(ca, cb)])), vals.append(np.array([np.sum(v)]))
            else:
                ca, cb = np.broadcast_to(ca, (N,)), np.broadcast_to(cb, (N,))
                rows.append(np.maximum(ca, cb)), cols.append(np.minimum(ca, cb)), vals.append(np.broadcast_to(v, (N,)).astype(float))

        # ---- Mayer term and terminal constraints (:277-298): functions of (xf, tf, x0, t0, a)
        one = np.ones(1)
        x0 = _V(Dual2.variable(X[0:1, s], ("x0", s)) * (1.0 / o.scale_x[s]) for s in range(nx))
        xf = _V(Dual2.variable(X[N - 1: N, s], ("xf", s)) * (1.0 / o.scale_x[s]) for s in range(nx))
        a1 = _V(Dual2.variable(A[m: m + 1], ("a", m)) * (1.0 / o.scale_a[m]) for m in range(na))
        t0s, tfs = Dual2.variable(T0 * one, ("T0",)) * (1.0 / st), Dual2.variable(TF * one, ("TF",)) * (1.0 / st)
        theta = Dual2(np.zeros(1))
        M = o.get_terminal_costs(ph)(xf, tfs, x0, t0s, a1)
        M = M[0] if isinstance(M, (list, tuple)) else M
        if isinstance(M, Dual2):
            theta = theta + M * lam_f
        if R["ntc"]:
            tc = _components(o.get_terminal_constraints(ph)(xf, tfs, x0, t0s, a1), R["ntc"])
            for r_ in range(R["ntc"]):
                if isinstance(tc[r_], Dual2):
                    theta = theta + tc[r_] * lam[base + R["TC"] + r_]

        def tcol(key):
            return {"x0": lambda: ora.colX(ph, 0, key[1]), "xf": lambda: ora.colX(ph, N - 1, key[1]),
                    "a": lambda: ora.colA(ph, key[1]), "T0": lambda: ora.colT0(ph), "TF": lambda: ora.colTF(ph)}[key[0]]()

        for (ka, kb), v in theta.H.items():
            ca, cb = tcol(ka), tcol(kb)
            rows.append(np.array([max(ca, cb)])), cols.append(np.array([min(ca, cb)])), vals.append(np.asarray(v, float).reshape(1))
        if hasattr(ora, "_extra_hessian"):  # constraint blocks a subclass appends to the phase (oracle/adaptive.py)
            ora._extra_hessian(ph, z, lam[base: base + R["n"]], rows, cols, vals)

    if not rows:
        return sp.csr_matrix((ora.n_z, ora.n_z))
    Hm = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ora.n_z, ora.n_z)).tocsr()
    Hm.sort_indices()
    return Hm
