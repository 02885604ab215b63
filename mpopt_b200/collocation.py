"""``mp.CollocationRoots`` / ``mp.Collocation`` surface, computed on the GPU.

Mirrors the reference classes (/root/reference/mpopt/mpopt.py:3706-4131 and :4134-4276): same
constructor arguments, attribute names (``roots``, ``tau0``, ``tau1``, ``poly_orders``) and method
names / argument meaning, so code written against ``mp.Collocation`` keeps working.  Every table value
comes from the CUDA kernels behind ``mpx_collocation_tables / _basis_at / _weights`` (csrc/mpx_tables.cuh);
composites are assembled on the host as ``scipy.sparse`` matrices instead of the reference's dense
(N-1) x N numpy arrays, which do not fit in memory at the sizes this package targets.

Differences, on purpose: the ``"numerical"`` ``D_MATRIX_METHOD`` (np.poly1d, inaccurate above p ~ 15) is not
offered -- both settings give the product-form values; quadrature weights are exact instead of IDAS-integrated.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib


def _tables(scheme, deg, tmin, tmax, device=0):
    n1 = deg + 1
    r, D, w, C = np.empty(n1), np.empty((n1, n1)), np.empty(n1), np.empty((deg, n1))
    _lib.check(_lib.lib().mpx_collocation_tables(_lib.SCHEMES[scheme], deg, float(tmin), float(tmax), device,
                                                 *[_lib.ptr(a) for a in (r, D, w, C)]))
    return r, D, w, C


class CollocationRoots:
    """Node sets LGR / LGL / CGL (mpopt.py:4134-4276).  ``_TAU_MIN/_TAU_MAX`` are read at construction."""

    _TAU_MIN = -1
    _TAU_MAX = 1

    def __init__(self, scheme: str = "LGR", device: int = 0):
        if scheme not in _lib.SCHEMES:
            raise ValueError(f"scheme must be one of {sorted(_lib.SCHEMES)}")
        self.scheme, self.device = scheme, device
        self._taus_fn = self.get_collocation_points(scheme, device)

    @classmethod
    def get_collocation_points(cls, scheme: str, device: int = 0):
        tmin, tmax = float(cls._TAU_MIN), float(cls._TAU_MAX)

        def taus(deg):
            if deg == 0:  # mpopt.py:4228-4229
                return np.array([0.0])
            return _tables(scheme, int(deg), tmin, tmax, device)[0]

        return taus

    @classmethod
    def roots_legendre_gauss_radau(cls, tau_min=-1, tau_max=1):
        return lambda deg: _tables("LGR", deg, tau_min, tau_max)[0]

    @classmethod
    def roots_legendre_gauss_lobatto(cls, tau_min=-1, tau_max=1):
        return lambda deg: _tables("LGL", deg, tau_min, tau_max)[0]

    @classmethod
    def roots_chebyshev_gauss_lobatto(cls, tau_min=-1, tau_max=1):
        return lambda deg: _tables("CGL", deg, tau_min, tau_max)[0]


class Collocation:
    D_MATRIX_METHOD = "symbolic"

    def __init__(self, poly_orders=(), scheme: str = "LGR", polynomial_type: str = "lagrange", device: int = 0):
        self.poly_orders = list(poly_orders)
        self.scheme, self.device = scheme, device
        cr = CollocationRoots(scheme, device)
        self._taus_fn = cr._taus_fn
        self.tau0, self.tau1 = float(cr._TAU_MIN), float(cr._TAU_MAX)  # mpopt.py:3741-3742
        self.roots, self._D, self._w, self._Cmid = {}, {}, {}, {}
        self.unique_polys = set(self.poly_orders)
        self.init_polynomials(self.unique_polys)

    def init_polynomials(self, poly_orders):
        for d in poly_orders:
            self.roots[d], self._D[d], self._w[d], self._Cmid[d] = _tables(self.scheme, int(d), self.tau0, self.tau1,
                                                                            self.device)

    def _basis_at(self, key, taus, order):
        taus = np.ascontiguousarray(np.atleast_1d(np.asarray(taus, dtype=float)))
        out = np.empty((len(taus), key + 1))
        if len(taus):
            _lib.check(_lib.lib().mpx_collocation_basis_at(_lib.SCHEMES[self.scheme], int(key), self.tau0, self.tau1,
                                                           self.device, order, len(taus), _lib.ptr(taus), _lib.ptr(out)))
        return out

    def get_diff_matrix(self, key, taus=None, order: int = 1):
        """D[i, j] = d^order l_j / dt^order at the nodes or at ``taus`` (mpopt.py:3815-3849)."""
        if key not in self.roots:
            self.init_polynomials([key])
        if taus is None and order == 1:
            return self._D[key].copy()
        return self._basis_at(key, self.roots[key] if taus is None else taus, order)

    def get_quadrature_weights(self, key, tau0=None, tau1=None):
        """w_j = int_{tau0}^{tau1} l_j (mpopt.py:3851-3882)."""
        if key not in self.roots:
            self.init_polynomials([key])
        if tau0 is None and tau1 is None:
            return self._w[key].copy()
        tau0 = self.tau0 if tau0 is None else tau0
        tau1 = self.tau1 if tau1 is None else tau1
        w = np.empty(key + 1)
        _lib.check(_lib.lib().mpx_collocation_weights(_lib.SCHEMES[self.scheme], int(key), self.tau0, self.tau1,
                                                      self.device, float(tau0), float(tau1), _lib.ptr(w)))
        return w

    def get_interpolation_matrix(self, taus, degree):
        """C[i, j] = l_j(taus[i]) (mpopt.py:3884-3905)."""
        if degree not in self.roots:
            self.init_polynomials([degree])
        return self._basis_at(degree, taus, 0)

    # ---- per-segment dictionaries (mpopt.py:3907-3985)
    def get_diff_matrices(self, poly_orders=None, order: int = 1):
        return {d: self.get_diff_matrix(d, order=order) for d in (self.unique_polys if poly_orders is None else set(poly_orders))}

    def get_quad_weight_matrices(self, keys=None, tau0=None, tau1=None):
        return {d: self.get_quadrature_weights(d, tau0, tau1) for d in (self.unique_polys if keys is None else set(keys))}

    def get_interpolation_matrices(self, taus, poly_orders=None):
        po = self.poly_orders if poly_orders is None else poly_orders
        return {i: self.get_interpolation_matrix(taus[i], d) for i, d in enumerate(po)}

    def get_interpolation_Dmatrices_at(self, taus, keys=None, order: int = 1):
        keys = self.poly_orders if keys is None else keys
        return {i: self.get_diff_matrix(k, taus=taus[i], order=order) for i, k in enumerate(keys)}

    # ---- composites (mpopt.py:4015-4131), sparse
    def get_composite_differentiation_matrix(self, poly_orders=None, order: int = 1):
        """Staircase: block 0 is the full D of segment 0, block k >= 1 contributes rows 1.. (the shared node's row
        comes from the earlier segment)."""
        po = self.poly_orders if poly_orders is None else list(poly_orders)
        D = self.get_diff_matrices(po, order=order)
        N = sum(po) + 1
        rows, cols, vals = [], [], []
        start = 0
        for k, p in enumerate(po):
            lo = 0 if k == 0 else 1
            ii, jj = np.meshgrid(np.arange(lo, p + 1), np.arange(p + 1), indexing="ij")
            rows.append(start + ii.ravel()), cols.append(start + jj.ravel()), vals.append(D[p][lo:, :].ravel())
            start += p
        return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()

    def get_composite_quadrature_weights(self, poly_orders=None, tau0=None, tau1=None):
        """[w0[0], w0[1:], w1[1:], ...] as a 1 x N array -- w[0] of later segments is dropped, like the reference."""
        po = self.poly_orders if poly_orders is None else list(poly_orders)
        W = self.get_quad_weight_matrices(po, tau0, tau1)
        return np.concatenate([[W[po[0]][0]]] + [W[p][1:] for p in po]).reshape(1, -1)

    def _composite_blocks(self, blocks, po):
        N = sum(po) + 1
        rows, cols, vals = [], [], []
        r0 = c0 = 0
        for i, p in enumerate(po):
            B = np.asarray(blocks[i])
            if B.shape[0]:
                ii, jj = np.meshgrid(np.arange(B.shape[0]), np.arange(p + 1), indexing="ij")
                rows.append(r0 + ii.ravel()), cols.append(c0 + jj.ravel()), vals.append(B.ravel())
            r0 += B.shape[0]
            c0 += p
        if not rows:
            return sp.csr_matrix((0, N))
        return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, N)).tocsr()

    def get_composite_interpolation_matrix(self, taus, poly_orders=None):
        po = self.poly_orders if poly_orders is None else list(poly_orders)
        return self._composite_blocks(self.get_interpolation_matrices(taus, po), po)

    def get_composite_interpolation_Dmatrix_at(self, taus, poly_orders=None, order: int = 1):
        po = self.poly_orders if poly_orders is None else list(poly_orders)
        return self._composite_blocks(self.get_interpolation_Dmatrices_at(taus, keys=po, order=order), po)
