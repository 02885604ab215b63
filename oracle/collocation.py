"""Oracle (TEST INFRASTRUCTURE): collocation nodes and D / W / C tables.

CPU restatement, float64 numpy + scipy, of the reference's ``CollocationRoots``
(/root/reference/mpopt/mpopt.py:4134-4276) and ``Collocation`` (:3706-4131) in
its default ``D_MATRIX_METHOD = "symbolic"`` mode, i.e. the *product form*
``l_j(t) = prod_{i != j} (t - r_i) / (r_j - r_i)`` (:3999-4004) differentiated
exactly (what ``ca.gradient`` does at :3834) -- not the ``np.poly1d`` monomial
mode, which loses accuracy above p ~ 15.

One deliberate deviation, documented as quirk Q2 in SURVEY.md: the reference
integrates l_j with SUNDIALS-IDAS at default tolerances (:3869-3877, ~1e-7
accurate); the oracle integrates exactly with a Gauss-Legendre rule.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.special

TAU_MIN = -1.0
TAU_MAX = 1.0


# --------------------------------------------------------------------------- roots
def roots(scheme: str, deg: int, tau_min: float = TAU_MIN, tau_max: float = TAU_MAX) -> np.ndarray:
    """Collocation nodes of one segment of polynomial degree ``deg`` (deg+1 nodes).

    mpopt.py:4157-4188 (dispatch), :4208-4231 LGR, :4234-4259 LGL, :4262-4276 CGL,
    :4191-4205 LG (unofficial), :4183-4188 fallback (equally spaced).
    """
    if scheme == "LGR":
        if deg > 1:
            r = scipy.special.roots_jacobi(deg - 1, 1.0, 0.0)[0]  # :4220
            r = np.append(np.append(-1, r), 1.0)
            return tau_min + (tau_max - tau_min) / 2 * (r + 1)  # :4224
        if deg == 1:
            return np.array([tau_min, tau_max], dtype=float)
        return np.array([0.0])
    if scheme == "LGL":
        if deg > 1:
            r = scipy.special.roots_jacobi(deg - 1, 1.0, 1.0)[0]  # :4246
            r = np.append(np.append(-1, r), 1.0)
            return tau_min + (tau_max - tau_min) / 2 * (r + 1)
        if deg == 1:
            return np.array([tau_min, tau_max], dtype=float)
        return np.array([0.0])
    if scheme == "CGL":
        r = np.array([np.cos(np.pi * j / deg) for j in range(deg + 1)])[::-1]  # :4271
        return tau_min + (tau_max - tau_min) / 2 * (r + 1)
    if scheme == "LG":
        r = np.polynomial.legendre.leggauss(deg - 1)[0]  # :4200
        r = np.append(-1, r)
        return tau_min + (tau_max - tau_min) / 2 * (r + 1)
    # :4183-4188  unknown scheme -> equally spaced, called with n_nodes = deg
    return np.linspace(tau_min, tau_max, deg) if deg > 1 else np.array([tau_min, tau_max], dtype=float)


# --------------------------------------------------------------------------- basis
def interpolation_matrix(r: np.ndarray, taus) -> np.ndarray:
    """C[i, j] = l_j(taus[i])  (mpopt.py:3884-3905, product form :3999-4004)."""
    r = np.asarray(r, dtype=float)
    taus = np.atleast_1d(np.asarray(taus, dtype=float))
    n = len(r)
    C = np.ones((len(taus), n))
    for j in range(n):
        for i in range(n):
            if i != j:
                C[:, j] *= (taus - r[i]) / (r[j] - r[i])
    return C


def diff_matrix(r: np.ndarray, taus=None, order: int = 1) -> np.ndarray:
    """D[i, j] = d^order/dt^order l_j (taus[i])  (mpopt.py:3815-3849).

    Differentiates the product form term by term (no division by (tau - r_i)),
    so it is valid at the nodes as well as between them.
    """
    r = np.asarray(r, dtype=float)
    taus = r if taus is None else np.atleast_1d(np.asarray(taus, dtype=float))
    n = len(r)
    D = np.zeros((len(taus), n))
    for j in range(n):
        others = [i for i in range(n) if i != j]
        fac = {i: (taus - r[i]) / (r[j] - r[i]) for i in others}
        inv = {i: 1.0 / (r[j] - r[i]) for i in others}
        if order == 1:
            for k in others:
                term = np.full(len(taus), inv[k])
                for i in others:
                    if i != k:
                        term = term * fac[i]
                D[:, j] += term
        elif order == 2:
            for k in others:
                for l in others:
                    if l == k:
                        continue
                    term = np.full(len(taus), inv[k] * inv[l])
                    for i in others:
                        if i != k and i != l:
                            term = term * fac[i]
                    D[:, j] += term
        else:
            raise ValueError("order must be 1 or 2")
    return D


def quadrature_weights(r: np.ndarray, tau0: float, tau1: float) -> np.ndarray:
    """w_j = int_{tau0}^{tau1} l_j  (mpopt.py:3851-3882), integrated exactly."""
    r = np.asarray(r, dtype=float)
    n = len(r)
    nq = n // 2 + 1  # exact for degree <= 2 nq - 1 >= n - 1
    xq, wq = np.polynomial.legendre.leggauss(nq)
    tq = tau0 + (tau1 - tau0) / 2 * (xq + 1)
    return (tau1 - tau0) / 2 * (wq @ interpolation_matrix(r, tq))


def mid_points(tau: np.ndarray) -> np.ndarray:
    """mpopt.py:350-352."""
    return np.array([(tau[i] + tau[i + 1]) / 2.0 for i in range(len(tau) - 1)])


# --------------------------------------------------------------------------- tables per degree
class Tables:
    """roots / D / w / C_mid for every unique degree of ``poly_orders`` (a1-a8)."""

    def __init__(self, poly_orders, scheme="LGR", tau_min=TAU_MIN, tau_max=TAU_MAX, override=None):
        """``override``: {degree: (roots, D, w, Cmid)} replaces the computed tables.  Used by the GPU parity
        tests to hand the oracle the device's own tables (already checked against these to 1e-11) so that
        the exact-zero folding of quirk Q10 -- which depends on rounding noise in analytically-zero
        entries such as the interior LGL diagonal -- is decided on identical numbers."""
        self.poly_orders = list(poly_orders)
        self.scheme = scheme
        self.tau0, self.tau1 = float(tau_min), float(tau_max)  # mpopt.py:3741-3742
        self.roots, self.D, self.w, self.Cmid = {}, {}, {}, {}
        for d in sorted(set(self.poly_orders)):
            if override is not None:
                self.roots[d], self.D[d], self.w[d], self.Cmid[d] = (np.array(a, dtype=float) for a in override[d])
                continue
            r = roots(scheme, d, tau_min, tau_max)
            self.roots[d] = r
            self.D[d] = diff_matrix(r)
            self.w[d] = quadrature_weights(r, self.tau0, self.tau1)
            self.Cmid[d] = interpolation_matrix(r, mid_points(r))  # mpopt.py:350-359

    # ---- composites (a9-a11)
    def composite_D(self, order: int = 1) -> sp.csr_matrix:
        """Block staircase of mpopt.py:4015-4039 as a sparse matrix.

        Every entry of a block is *assigned* (structurally present, including
        numerically zero values); callers that restate CasADi's SX folding drop
        the exact zeros themselves (quirk Q10).
        """
        po = self.poly_orders
        N = sum(po) + 1
        rows, cols, vals = [], [], []
        start = 0
        for k, p in enumerate(po):
            Dk = self.D[p] if order == 1 else diff_matrix(self.roots[p], order=order)
            if k == 0:
                ii, jj = np.meshgrid(np.arange(p + 1), np.arange(p + 1), indexing="ij")
                rows.append(ii.ravel()), cols.append(jj.ravel()), vals.append(Dk.ravel())
            else:
                ii, jj = np.meshgrid(np.arange(1, p + 1), np.arange(p + 1), indexing="ij")
                rows.append(start + ii.ravel()), cols.append(start + jj.ravel())
                vals.append(Dk[1:, :].ravel())
            start += p
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
        return A.tocsr()

    def composite_W(self) -> np.ndarray:
        """mpopt.py:4041-4064: [w_p0[0], w_p0[1:], w_p1[1:], ...] -- w[0] of segments k>=1 dropped (Q1)."""
        po = self.poly_orders
        return np.concatenate([[self.w[po[0]][0]]] + [self.w[p][1:] for p in po])

    def composite_interpolation(self, taus, D_order: int = 0) -> sp.csr_matrix:
        """mpopt.py:4066-4096 (D_order=0) and :4098-4131 (D_order=1|2), sparse instead of dense."""
        po = self.poly_orders
        N = sum(po) + 1
        rows, cols, vals = [], [], []
        r0 = c0 = 0
        cache = {}
        for k, p in enumerate(po):
            t = np.asarray(taus[k], dtype=float)
            if len(t):
                key = (p, t.tobytes())
                if key not in cache:  # same degree, same points -> same block
                    cache[key] = (interpolation_matrix(self.roots[p], t) if D_order == 0
                                  else diff_matrix(self.roots[p], t, D_order))
                blk = cache[key]
                ii, jj = np.meshgrid(np.arange(len(t)), np.arange(p + 1), indexing="ij")
                rows.append(r0 + ii.ravel()), cols.append(c0 + jj.ravel()), vals.append(blk.ravel())
            r0 += len(t)
            c0 += p
        if not rows:
            return sp.csr_matrix((0, N))
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, N))
        return A.tocsr()

    def composite_mid_interpolation(self) -> sp.csr_matrix:
        """The (N-1) x N matrix of mpopt.py:353-359."""
        return self.composite_interpolation([mid_points(self.roots[p]) for p in self.poly_orders])

    def composite_slope_continuity(self) -> sp.csr_matrix:
        """(K-1) x N matrix of mpopt.py:398-403: row k = D_k(tau1) - D_{k+1}(tau0)."""
        K = len(self.poly_orders)
        ends = [np.array([self.tau0, self.tau1]) for _ in range(K)]
        M = self.composite_interpolation(ends, D_order=1).tocsr()
        # [1:-1][::2] - [2:-1][::2]
        a = M[1:-1][::2]
        b = M[2:-1][::2]
        return (a - b).tocsr()
