"""Type-agnostic stand-in for the handful of ``casadi`` functions user callables use.

The reference's example problems call ``ca.sqrt / ca.exp / ca.sin / ca.acos /
ca.vertcat`` inside dynamics and constraint lambdas
(/root/reference/examples/singlephase/robot_arm.py:45-54,
/root/reference/examples/Multi-phase/multistage_launch_vehicle.py:70-91).  CasADi is
not a dependency of this package; ``from mpopt_b200 import ca`` gives the same names.
Every function dispatches on its argument by duck typing: an object with a method
of that name (the tracer's ``Expr``, any dual-number type) handles itself, plain
numbers and numpy arrays go to numpy.  No differentiation logic lives here.
"""
from __future__ import annotations

import numpy as np

pi = np.pi
inf = np.inf


class Vec(list):
    """Column vector of scalars with element-wise arithmetic (what ``ca.vertcat`` returns)."""

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return Vec(r) if isinstance(i, slice) else r

    def _zip(self, o, fn):
        if isinstance(o, (list, tuple)) or (isinstance(o, np.ndarray) and o.ndim > 0):
            o = list(o)
            if len(o) != len(self):
                raise ValueError("Vec: length mismatch")
            return Vec(fn(a, b) for a, b in zip(self, o))
        return Vec(fn(a, o) for a in self)

    def __add__(self, o): return self._zip(o, lambda a, b: a + b)
    def __radd__(self, o): return self._zip(o, lambda a, b: b + a)
    def __sub__(self, o): return self._zip(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._zip(o, lambda a, b: b - a)
    def __mul__(self, o): return self._zip(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._zip(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._zip(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._zip(o, lambda a, b: b / a)
    def __pow__(self, o): return self._zip(o, lambda a, b: a ** b)
    def __neg__(self): return Vec(-a for a in self)
    __array_priority__ = 2000.0
    __array_ufunc__ = None  # numpy scalars defer to the reflected operators above


def _flat(args):
    for a in args:
        if isinstance(a, (list, tuple)) or (isinstance(a, np.ndarray) and a.ndim > 0):
            yield from _flat(list(a))
        else:
            yield a


def vertcat(*args):
    return Vec(_flat(args))


horzcat = vertcat


def _unary(name, np_name=None):
    np_fn = getattr(np, np_name or name)

    def fn(x):
        if isinstance(x, (list, tuple)):
            return Vec(fn(e) for e in x)
        m = getattr(x, name, None)
        if m is not None and not isinstance(x, np.ndarray):
            return m()
        return np_fn(x)

    fn.__name__ = name
    return fn


sqrt = _unary("sqrt")
exp = _unary("exp")
log = _unary("log")
sin = _unary("sin")
cos = _unary("cos")
tan = _unary("tan")
asin = _unary("asin", "arcsin")
acos = _unary("acos", "arccos")
atan = _unary("atan", "arctan")
sinh = _unary("sinh")
cosh = _unary("cosh")
tanh = _unary("tanh")
fabs = _unary("fabs")
arcsin, arccos, arctan = asin, acos, atan


def sumsqr(v):
    return sum(e * e for e in _flat([v]))


def dot(a, b):
    return sum(x * y for x, y in zip(_flat([a]), _flat([b])))


def norm_2(v):
    return sqrt(sumsqr(v))


def sum1(v):
    return sum(_flat([v]))


def power(x, y):
    return x ** y
