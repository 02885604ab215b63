"""Oracle (TEST INFRASTRUCTURE): vectorised second-order dual numbers.

A ``Dual2`` carries a value array, first derivatives ``g[key]`` and second derivatives ``H[(k1, k2)]`` (k1 <= k2 in
sorted order) with respect to named inputs; a key is present iff the derivative is *structurally* non-zero, which is how
the oracle obtains the sparsity CasADi's symbolic Hessian would have.  Independent of the tracer (mpopt_b200/trace.py)
and of the first-order ``oracle/dual.py``; used by oracle/hessian.py to restate nlp_hess_l
(/root/reference/mpopt/mpopt.py:757).
"""
from __future__ import annotations

import numpy as np


def _pair(a, b):
    return (a, b) if a <= b else (b, a)


def _is_num(v):
    return isinstance(v, (int, float, np.integer, np.floating)) or (isinstance(v, np.ndarray) and v.dtype != object)


class Dual2:
    __array_priority__ = 3000.0

    def __init__(self, val, g=None, H=None):
        self.val = np.asarray(val, dtype=float)
        self.g = dict(g or {})
        self.H = dict(H or {})

    @staticmethod
    def variable(val, key):
        val = np.asarray(val, dtype=float)
        return Dual2(val, {key: np.ones_like(val)})

    # -- composition rule: y = f(u):  g_y = f' g_u ;  H_y = f' H_u + f'' g_u g_u^T   (f2 None: f'' == 0 structurally)
    def _chain(self, val, f1, f2):
        g = {k: f1 * d for k, d in self.g.items()}
        H = {k: f1 * d for k, d in self.H.items()}
        if f2 is not None:
            keys = sorted(self.g)
            for i, a in enumerate(keys):
                for b in keys[i:]:
                    t = f2 * self.g[a] * self.g[b]
                    H[(a, b)] = H[(a, b)] + t if (a, b) in H else t
        return Dual2(val, g, H)

    # -- arithmetic
    def __neg__(self):
        return self._chain(-self.val, -1.0, None)

    def __pos__(self):
        return self

    def __add__(self, o):
        if isinstance(o, list):  # scalar (op) vector: handled element-wise by the vector's reflected operator
            return NotImplemented
        return _add(self, o, 1.0)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, list):
            return NotImplemented
        return _add(self, o, -1.0)

    def __rsub__(self, o):
        return _add(-self, o, 1.0)

    def __mul__(self, o):
        if isinstance(o, list):
            return NotImplemented
        return _mul(self, o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, list):
            return NotImplemented
        if isinstance(o, Dual2):
            return _mul(self, o._chain(1.0 / o.val, -1.0 / o.val ** 2, 2.0 / o.val ** 3))
        return self._chain(self.val / o, 1.0 / np.asarray(o, dtype=float), None)

    def __rtruediv__(self, o):
        return _mul(self._chain(1.0 / self.val, -1.0 / self.val ** 2, 2.0 / self.val ** 3), o)

    def __pow__(self, o):
        if isinstance(o, Dual2):
            return (o * self.log()).exp()
        n = float(o)
        if n == 1.0:
            return self
        if n == 2.0:
            return self._chain(self.val ** 2, 2.0 * self.val, 2.0 * np.ones_like(self.val))
        return self._chain(self.val ** n, n * self.val ** (n - 1.0), n * (n - 1.0) * self.val ** (n - 2.0))

    def __rpow__(self, o):
        return (self * np.log(float(o))).exp()

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        name = {"arccos": "acos", "arcsin": "asin", "arctan": "atan", "absolute": "fabs"}.get(ufunc.__name__, ufunc.__name__)
        two = {"add": lambda a, b: a + b, "subtract": lambda a, b: a - b, "multiply": lambda a, b: a * b,
               "true_divide": lambda a, b: a / b, "divide": lambda a, b: a / b, "power": lambda a, b: a ** b}
        if name in two:
            a, b = inputs
            if not isinstance(a, Dual2):  # ndarray (op) Dual2 -> reflected operator of the Dual2
                return {"add": b.__radd__, "subtract": b.__rsub__, "multiply": b.__rmul__, "true_divide": b.__rtruediv__,
                        "divide": b.__rtruediv__, "power": b.__rpow__}[name](a)
            return two[name](a, b)
        if name == "negative":
            return -inputs[0]
        if name == "square":
            return inputs[0] ** 2
        if len(inputs) == 1 and hasattr(self, name):
            return getattr(self, name)()
        return NotImplemented

    # -- elementary functions (same names as the casadi-style shim, which dispatches by duck typing)
    def sqrt(self):
        r = np.sqrt(self.val)
        return self._chain(r, 0.5 / r, -0.25 / (r * self.val))

    def exp(self):
        r = np.exp(self.val)
        return self._chain(r, r, r)

    def log(self):
        return self._chain(np.log(self.val), 1.0 / self.val, -1.0 / self.val ** 2)

    def sin(self):
        s, c = np.sin(self.val), np.cos(self.val)
        return self._chain(s, c, -s)

    def cos(self):
        s, c = np.sin(self.val), np.cos(self.val)
        return self._chain(c, -s, -c)

    def tan(self):
        t = np.tan(self.val)
        return self._chain(t, 1.0 + t * t, 2.0 * t * (1.0 + t * t))

    def asin(self):
        q = 1.0 - self.val ** 2
        return self._chain(np.arcsin(self.val), 1.0 / np.sqrt(q), self.val / q ** 1.5)

    def acos(self):
        q = 1.0 - self.val ** 2
        return self._chain(np.arccos(self.val), -1.0 / np.sqrt(q), -self.val / q ** 1.5)

    def atan(self):
        q = 1.0 + self.val ** 2
        return self._chain(np.arctan(self.val), 1.0 / q, -2.0 * self.val / q ** 2)

    def sinh(self):
        return self._chain(np.sinh(self.val), np.cosh(self.val), np.sinh(self.val))

    def cosh(self):
        return self._chain(np.cosh(self.val), np.sinh(self.val), np.cosh(self.val))

    def tanh(self):
        t = np.tanh(self.val)
        return self._chain(t, 1.0 - t * t, -2.0 * t * (1.0 - t * t))

    def fabs(self):
        return self._chain(np.abs(self.val), np.sign(self.val), None)

    arccos, arcsin, arctan = acos, asin, atan
    __abs__ = fabs


def _add(a, b, sign):
    if not isinstance(b, Dual2):
        return Dual2(a.val + sign * np.asarray(b, dtype=float), a.g, a.H)
    g = dict(a.g)
    for k, d in b.g.items():
        g[k] = g[k] + sign * d if k in g else sign * d
    H = dict(a.H)
    for k, d in b.H.items():
        H[k] = H[k] + sign * d if k in H else sign * d
    return Dual2(a.val + sign * b.val, g, H)


def _mul(a, b):
    if not isinstance(b, Dual2):
        c = np.asarray(b, dtype=float)
        if c.ndim == 0 and float(c) == 0.0:
            return 0.0  # SX folds 0 * x -> 0: a plain constant, so that later products with it fold too
        return a._chain(a.val * c, c, None)
    g = {k: b.val * d for k, d in a.g.items()}
    for k, d in b.g.items():
        g[k] = g[k] + a.val * d if k in g else a.val * d
    H = {k: b.val * d for k, d in a.H.items()}
    for k, d in b.H.items():
        H[k] = H[k] + a.val * d if k in H else a.val * d
    for ka, da in a.g.items():
        for kb, db in b.g.items():
            k = _pair(ka, kb)
            t = da * db * (2.0 if ka == kb else 1.0)
            H[k] = H[k] + t if k in H else t
    return Dual2(a.val * b.val, g, H)
