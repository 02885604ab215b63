"""TEST INFRASTRUCTURE -- run the UNMODIFIED reference (/root/reference/mpopt/mpopt.py) in this container.

The reference needs CasADi and matplotlib at import time; neither is installable here.  ``oracle/refrun/stubs`` holds a
small expression engine with the slice of CasADi's API the reference calls (see its docstring) and an empty matplotlib.
With those on ``sys.path`` the reference's own classes -- ``OCP``, ``Collocation``, ``CollocationRoots``, ``mpopt``,
``mpopt_adaptive`` -- import and run as they are: ``create_nlp()`` (mpopt.py:584-644) builds the reference's NLP
``{f, x, g, p}`` out of the reference's own formulas, and this module evaluates it and its first derivatives at a point.

That is what pins the VALUES of ``g``, ``jac_g``, ``f``, ``grad_f``, the Jacobian pattern, the bounds and the initial
guess of the oracle (and through it of the CUDA path) to the reference itself rather than to a restatement:
``tests/golden/make_reference_golden.py`` stores the results as ``tests/golden/ref_*.npz`` (the reference does not exist
on the GPU box), ``tests/test_reference_golden.py`` compares.

Only importable where /root/reference exists.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MPOPT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "mpopt", "mpopt.py"))


def load_reference():
    """The reference module, imported from where it lies, with the stand-ins for its two missing dependencies."""
    if not available():
        raise RuntimeError("the reference tree is not present on this machine")
    sys.dont_write_bytecode = True  # the reference tree is read-only by agreement: no __pycache__ next to its sources
    for p in (os.path.join(HERE, "stubs"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import casadi  # noqa: F401  (must resolve to the stand-in)

    if not casadi.__file__.startswith(HERE):
        raise RuntimeError("a real casadi is importable: use it instead of the stand-in")
    import mpopt.mpopt as ref

    return ref


def reference_ocp(ref, factory):
    """Build one of ``mpopt_b200.problems``' OCPs as an instance of the REFERENCE's ``OCP`` class: the problem
    definitions only set attributes and callables, so the class is swapped for the duration of the call."""
    import mpopt_b200.problems as problems

    saved = problems.OCP
    problems.OCP = ref.OCP
    try:
        return factory()
    finally:
        problems.OCP = saved


class ReferenceNLP:
    """The NLP of ``ref.mpopt(ocp, n_segments, poly_orders, scheme)`` (or ``ref.mpopt_adaptive``), evaluated numerically."""

    def __init__(self, ref, ocp, n_segments, poly_orders, scheme, adaptive=False):
        import casadi as ca

        cls = ref.mpopt_adaptive if adaptive else ref.mpopt
        self.mpo = cls(ocp, n_segments, poly_orders, scheme)
        self.mpo._MUTE_ = True
        self.nlp, self.bounds = self.mpo.create_nlp()
        x, p = self.nlp["x"], self.nlp["p"]
        self.n_z, self.n_p = x.shape[0], p.shape[0] * p.shape[1]
        self.n_g = self.nlp["g"].shape[0]
        # mpopt_adaptive keeps the widths inside x and drops "p" before it builds its solver (mpopt.py:3191-3193)
        in_x = {id(e) for e in x.a.flat}
        self.p_in_x = self.n_p > 0 and all(id(e) in in_x for e in p.a.flat)
        if self.p_in_x:
            self.n_p = 0
        self._fn = ca.Function("nlp", [x] if self.p_in_x else [x, p], [self.nlp["g"], self.nlp["f"]])

    def evaluate(self, z, p):
        """(f, g, grad_f, jac_g as CSR) of the reference's NLP at ``(z, p)``."""
        args = [np.asarray(z, float)] if self.p_in_x else [np.asarray(z, float), np.asarray(p, float)]
        (g, dg), (f, df) = self._fn.forward_sparse(args, wrt=0)
        indptr, indices, data = [0], [], []
        for row in dg:
            cols = sorted(row)
            indices.extend(cols)
            data.extend(row[c] for c in cols)
            indptr.append(len(indices))
        J = sp.csr_matrix((np.array(data, float), np.array(indices, np.int64), np.array(indptr, np.int64)),
                          shape=(self.n_g, self.n_z))
        grad = np.zeros(self.n_z)
        for c, v in df[0].items():
            grad[c] = v
        return float(f[0]), g, grad, J

    def hess_l(self, z, p, lam_f, lam_g):
        """Lower triangle (CSR) of the Hessian of ``lam_f f + lam_g . g`` -- the Lagrangian ``ca.nlpsol`` differentiates
        (mpopt.py:757) -- of the reference's NLP at ``(z, p)``; structural entries of every row with a non-zero role."""
        args = [np.asarray(z, float)] if self.p_in_x else [np.asarray(z, float), np.asarray(p, float)]
        (_, _, hg), (_, _, hf) = self._fn.forward2_sparse(args, wrt=0)
        H = {}
        for lam, rows in ((np.asarray(lam_g, float), hg), (np.array([float(lam_f)]), hf)):
            for l, h in zip(lam, rows):
                for k, v in h.items():
                    H[k] = H.get(k, 0.0) + l * v
        keys = sorted(H)
        return sp.csr_matrix((np.array([H[k] for k in keys], float), (np.array([k[0] for k in keys], np.int64),
                                                                       np.array([k[1] for k in keys], np.int64))),
                             shape=(self.n_z, self.n_z))

    def initial_guess(self):
        return np.asarray(self.mpo.initialize_solution(), float).ravel()

    def all_bounds(self):
        b = self.bounds
        return tuple(np.asarray(b[k], float).ravel() for k in ("lbx", "ubx", "lbg", "ubg"))
