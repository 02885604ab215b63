"""Oracle (TEST INFRASTRUCTURE): vectorised forward-mode dual numbers.

Restates, on the CPU, what CasADi's SX layer does for the reference at
/root/reference/mpopt/mpopt.py:757 (``ca.nlpsol`` derives ``jac_g``/``grad_f`` by
algorithmic differentiation with *structural* sparsity): every ``Dual`` carries a
value array (one entry per collocation node) and a dict ``{variable: d/dvariable}``
whose key set is the structural dependency set.  CasADi (casadi==3.6.0, absent
here) simplifies SX expressions on construction -- ``0*x -> 0``, ``x+0 -> x``,
``1*x -> x``, ``x-x -> 0`` -- which changes the Jacobian pattern; the same
folds are applied here.

Independent of the product tracer in ``mpopt_b200/trace.py`` by construction: no
imports from the package.
"""
from __future__ import annotations

import math
import numbers

import numpy as np


def _is_num(v):
    return isinstance(v, (numbers.Real, np.floating, np.integer)) or (
        isinstance(v, np.ndarray) and v.dtype != object and v.ndim == 0
    )


def _unwrap(v):
    """1-element numeric arrays / lists behave as scalars (ocp.t00[phase] is shape (1,))."""
    if isinstance(v, np.ndarray) and v.dtype != object and v.size == 1:
        return float(v.reshape(-1)[0])
    return v


class Dual:
    __array_priority__ = 1000.0

    def __init__(self, val, der=None):
        self.val = np.asarray(val, dtype=float)
        self.der = dict(der or {})

    # -- numpy interop: ufuncs on a Dual dispatch to the methods below
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        name = ufunc.__name__
        table = {
            "add": lambda a, b: _add(a, b),
            "subtract": lambda a, b: _sub(a, b),
            "multiply": lambda a, b: _mul(a, b),
            "true_divide": lambda a, b: _div(a, b),
            "divide": lambda a, b: _div(a, b),
            "power": lambda a, b: _pow(a, b),
            "negative": lambda a: -a,
            "positive": lambda a: a,
            "absolute": lambda a: a.fabs(),
            "square": lambda a: _mul(a, a),
        }
        if name in table:
            return table[name](*inputs)
        alias = {"arccos": "acos", "arcsin": "asin", "arctan": "atan"}
        name = alias.get(name, name)
        if len(inputs) == 1 and hasattr(self, name):
            return getattr(self, name)()
        return NotImplemented

    # -- arithmetic
    def __add__(self, o):
        if isinstance(o, list):  # scalar (op) vector: the vector's reflected operator applies it element-wise
            return NotImplemented
        return _add(self, o)

    def __radd__(self, o):
        return _add(o, self)

    def __sub__(self, o):
        if isinstance(o, list):  # scalar (op) vector: the vector's reflected operator applies it element-wise
            return NotImplemented
        return _sub(self, o)

    def __rsub__(self, o):
        return _sub(o, self)

    def __mul__(self, o):
        if isinstance(o, list):  # scalar (op) vector: the vector's reflected operator applies it element-wise
            return NotImplemented
        return _mul(self, o)

    def __rmul__(self, o):
        return _mul(o, self)

    def __truediv__(self, o):
        if isinstance(o, list):  # scalar (op) vector: the vector's reflected operator applies it element-wise
            return NotImplemented
        return _div(self, o)

    def __rtruediv__(self, o):
        return _div(o, self)

    def __pow__(self, o):
        return _pow(self, o)

    def __rpow__(self, o):
        return _pow(o, self)

    def __neg__(self):
        return Dual(-self.val, {k: -d for k, d in self.der.items()})

    def __pos__(self):
        return self

    # -- elementary functions (names shared with the casadi-style shim by duck typing)
    def _chain(self, val, dval):
        return Dual(val, {k: dval * d for k, d in self.der.items()})

    def sqrt(self):
        r = np.sqrt(self.val)
        return self._chain(r, 0.5 / r)

    def exp(self):
        r = np.exp(self.val)
        return self._chain(r, r)

    def log(self):
        return self._chain(np.log(self.val), 1.0 / self.val)

    def sin(self):
        return self._chain(np.sin(self.val), np.cos(self.val))

    def cos(self):
        return self._chain(np.cos(self.val), -np.sin(self.val))

    def tan(self):
        t = np.tan(self.val)
        return self._chain(t, 1.0 + t * t)

    def asin(self):
        return self._chain(np.arcsin(self.val), 1.0 / np.sqrt(1.0 - self.val**2))

    def acos(self):
        return self._chain(np.arccos(self.val), -1.0 / np.sqrt(1.0 - self.val**2))

    def atan(self):
        return self._chain(np.arctan(self.val), 1.0 / (1.0 + self.val**2))

    def sinh(self):
        return self._chain(np.sinh(self.val), np.cosh(self.val))

    def cosh(self):
        return self._chain(np.cosh(self.val), np.sinh(self.val))

    def tanh(self):
        t = np.tanh(self.val)
        return self._chain(t, 1.0 - t * t)

    def fabs(self):
        return self._chain(np.abs(self.val), np.sign(self.val))

    arccos, arcsin, arctan = acos, asin, atan
    __abs__ = fabs


def _lift(v):
    v = _unwrap(v)
    if isinstance(v, Dual):
        return v, None
    if _is_num(v):
        return None, float(v)
    raise TypeError(f"unsupported operand for Dual arithmetic: {type(v)!r}")


def _add(a, b):
    da, ca_ = _lift(a)
    db, cb = _lift(b)
    if da is None:
        return db if ca_ == 0.0 else Dual(ca_ + db.val, db.der)
    if db is None:
        return da if cb == 0.0 else Dual(da.val + cb, da.der)
    der = dict(da.der)
    for k, d in db.der.items():
        der[k] = der[k] + d if k in der else d
    return Dual(da.val + db.val, der)


def _sub(a, b):
    da, ca_ = _lift(a)
    db, cb = _lift(b)
    if da is not None and da is db:
        return 0.0  # x - x -> 0
    if da is None:
        return -db if ca_ == 0.0 else Dual(ca_ - db.val, {k: -d for k, d in db.der.items()})
    if db is None:
        return da if cb == 0.0 else Dual(da.val - cb, da.der)
    der = dict(da.der)
    for k, d in db.der.items():
        der[k] = der[k] - d if k in der else -d
    return Dual(da.val - db.val, der)


def _mul(a, b):
    da, ca_ = _lift(a)
    db, cb = _lift(b)
    if da is None:
        da, ca_, db, cb = db, cb, da, ca_
    if db is None:  # Dual * const
        if cb == 0.0:
            return 0.0
        if cb == 1.0:
            return da
        return Dual(da.val * cb, {k: d * cb for k, d in da.der.items()})
    der = {k: d * db.val for k, d in da.der.items()}
    for k, d in db.der.items():
        t = d * da.val
        der[k] = der[k] + t if k in der else t
    return Dual(da.val * db.val, der)


def _div(a, b):
    da, ca_ = _lift(a)
    db, cb = _lift(b)
    if db is None:  # Dual / const
        if cb == 1.0:
            return da
        return Dual(da.val / cb, {k: d / cb for k, d in da.der.items()})
    if da is None:  # const / Dual
        if ca_ == 0.0:
            return 0.0
        q = ca_ / db.val
        return Dual(q, {k: -q / db.val * d for k, d in db.der.items()})
    if da is db:
        return 1.0
    q = da.val / db.val
    der = {k: d / db.val for k, d in da.der.items()}
    for k, d in db.der.items():
        t = -q / db.val * d
        der[k] = der[k] + t if k in der else t
    return Dual(q, der)


def _pow(a, b):
    da, ca_ = _lift(a)
    db, cb = _lift(b)
    if db is None:  # Dual ** const
        if cb == 0.0:
            return 1.0
        if cb == 1.0:
            return da
        if cb == 2.0:
            return _mul(da, da)
        return da._chain(da.val**cb, cb * da.val ** (cb - 1.0))
    if da is None:  # const ** Dual
        r = ca_**db.val
        return db._chain(r, r * math.log(ca_))
    r = da.val**db.val
    der = {k: db.val * da.val ** (db.val - 1.0) * d for k, d in da.der.items()}
    for k, d in db.der.items():
        t = r * np.log(da.val) * d
        der[k] = der[k] + t if k in der else t
    return Dual(r, der)


def flatten(out):
    """Flatten what a user callable may return (scalar, list, nested list, object array) to a list."""
    if out is None:
        return None
    if isinstance(out, Dual) or _is_num(out):
        return [out]
    res = []
    for o in (out.ravel().tolist() if isinstance(out, np.ndarray) else out):
        res.extend(flatten(o))
    return res


class Vec(list):
    """List with element-wise arithmetic and slice -> Vec, standing in for the SX column
    vectors the reference hands to user callables (mpopt.py:196-197): ``x[:3]``,
    ``scalar * x[3:6]`` and ``x[-1]`` all occur in /root/reference/examples."""

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return Vec(r) if isinstance(i, slice) else r

    def _bin(self, o, fn):
        if isinstance(o, (list, tuple, np.ndarray)) and not _is_num(o):
            o = list(o)
            assert len(o) == len(self)
            return Vec(fn(a, b) for a, b in zip(self, o))
        return Vec(fn(a, o) for a in self)

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._bin(o, lambda a, b: b + a)

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._bin(o, lambda a, b: b * a)

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._bin(o, lambda a, b: b / a)

    def __neg__(self):
        return Vec(-a for a in self)

    def __pow__(self, o):
        return self._bin(o, lambda a, b: a**b)
