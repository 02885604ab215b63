"""TEST INFRASTRUCTURE -- a small scalar-expression engine exposing the slice of CasADi's Python API that
/root/reference/mpopt/mpopt.py calls, so that the REFERENCE'S OWN transcription code can be imported and run in this
container (CasADi itself is not installable here).  Nothing under mpopt_b200/ imports it.

What it is: ``SX`` / ``DM`` dense matrices of scalar expression nodes with CasADi's indexing conventions (column-major
``M[:]``, linear ``M[k]``, 2-D slices), ``vertcat / mtimes / kron / diag / solve / sum1 / gradient / jacobian``,
``Function`` (numeric evaluation of an expression graph), an exact stand-in for ``integrator`` (polynomial ODE right-hand
sides only -- the reference integrates Lagrange polynomials with it, mpopt.py:3869-3877) and a ``nlpsol`` that records the
NLP without solving it.  Scalar operations apply the construction-time simplifications of CasADi's ``SXElem::binary``
(``0*x -> 0``, ``x+0 -> x``, ``x-x -> 0``, ``1*x -> x``, ``x*x -> sq(x)`` ...), because the reference's Jacobian pattern
depends on them (exact-zero table entries disappear from the pattern).

Which CasADi: the reference pins casadi==3.6.0 (requirements.txt:4; >=3.5.5 in setup.py:29); it is third-party, un-vendored
and absent from /root/reference and from this image.  The call sites served here are mpopt.py:114-128, 152, 177-206,
228-232, 255-298, 317-321, 360-372, 406-408, 455-458, 484-516, 537-543, 624-627, 757, 804, 873-896, 996-1076, 1307-1344,
1464-1482, 1512-1573, 3038-3131, 3206, 3243-3268, 3718, 3830-3903, 4000-4062.

What it is not: CasADi.  Values produced through it are the reference's FORMULAS evaluated in IEEE double arithmetic; the
order of floating-point operations inside ``mtimes`` follows CasADi's (ascending inner index), reverse-mode AD is replaced
by forward-mode (same numbers up to rounding).  Jacobian sparsity = structural dependence of the simplified graph, which
is what CasADi's sparsity propagation computes.
"""
from __future__ import annotations

import math

import numpy as np

pi = math.pi
inf = math.inf

# --------------------------------------------------------------------------------------------------------------------
# scalar nodes


class E:
    __slots__ = ("op", "a", "b", "v")

    def __init__(self, op, a=None, b=None, v=None):
        self.op, self.a, self.b, self.v = op, a, b, v

    def __repr__(self):
        if self.op == "c":
            return repr(self.v)
        if self.op == "s":
            return str(self.v)
        if self.b is None:
            return f"{self.op}({self.a!r})"
        return f"{self.op}({self.a!r},{self.b!r})"


def C(v):
    return E("c", v=float(v))


ZERO, ONE = C(0.0), C(1.0)


def _isc(x, v=None):
    return x.op == "c" and (v is None or x.v == v)


def _eq(x, y, depth=1):
    """SXElem::is_equal with CasADi's default depth 1."""
    if x is y:
        return True
    if x.op == "c" and y.op == "c":
        return x.v == y.v
    if depth > 0 and x.op == y.op and x.op not in ("c", "s"):
        if x.b is None:
            return _eq(x.a, y.a, depth - 1)
        if _eq(x.a, y.a, depth - 1) and _eq(x.b, y.b, depth - 1):
            return True
        if x.op in ("add", "mul") and _eq(x.a, y.b, depth - 1) and _eq(x.b, y.a, depth - 1):
            return True
    return False


_UN = {
    "neg": lambda v: -v, "sq": lambda v: v * v, "sqrt": math.sqrt, "sin": math.sin, "cos": math.cos, "tan": math.tan,
    "exp": math.exp, "log": math.log, "asin": math.asin, "acos": math.acos, "atan": math.atan, "sinh": math.sinh,
    "cosh": math.cosh, "tanh": math.tanh, "fabs": abs, "inv": lambda v: 1.0 / v,
}


def _safe(fn, *v):
    try:
        return fn(*v)
    except (ValueError, OverflowError, ZeroDivisionError):
        return math.nan


def neg(x):
    if x.op == "c":
        return C(-x.v)
    if x.op == "neg":
        return x.a
    return E("neg", x)


def add(x, y):
    if x.op == "c" and y.op == "c":
        return C(x.v + y.v)
    if _isc(x, 0.0):
        return y
    if _isc(y, 0.0):
        return x
    if x.op == "neg":
        return sub(y, x.a)
    if y.op == "neg":
        return sub(x, y.a)
    if x.op == "sub" and _eq(x.b, y):
        return x.a
    if y.op == "sub" and _eq(x, y.b):
        return y.a
    return E("add", x, y)


def sub(x, y):
    if x.op == "c" and y.op == "c":
        return C(x.v - y.v)
    if _isc(y, 0.0):
        return x
    if _isc(x, 0.0):
        return neg(y)
    if _eq(x, y):
        return ZERO
    if y.op == "neg":
        return add(x, y.a)
    if x.op == "add" and _eq(x.b, y):
        return x.a
    if x.op == "add" and _eq(x.a, y):
        return x.b
    if y.op == "add" and _eq(x, y.b):
        return neg(y.a)
    if y.op == "add" and _eq(x, y.a):
        return neg(y.b)
    if x.op == "neg":
        return neg(add(x.a, y))
    return E("sub", x, y)


def mul(x, y):
    if x.op == "c" and y.op == "c":
        return C(x.v * y.v)
    if _eq(x, y):
        return unary("sq", x)
    if x.op != "c" and y.op == "c":
        return mul(y, x)
    if _isc(x, 0.0) or _isc(y, 0.0):
        return ZERO
    if _isc(x, 1.0):
        return y
    if _isc(y, 1.0):
        return x
    if _isc(y, -1.0):
        return neg(x)
    if _isc(x, -1.0):
        return neg(y)
    if x.op == "inv":
        return div(y, x.a)
    if y.op == "inv":
        return div(x, y.a)
    if x.op == "c" and y.op == "mul" and y.a.op == "c" and x.v * y.a.v == 1.0:
        return y.b
    if x.op == "c" and y.op == "div" and y.b.op == "c" and x.v == y.b.v:
        return y.a
    if x.op == "neg":
        return neg(mul(x.a, y))
    if y.op == "neg":
        return neg(mul(x, y.a))
    return E("mul", x, y)


def div(x, y):
    if x.op == "c" and y.op == "c":
        return C(_safe(lambda a, b: a / b, x.v, y.v) if y.v != 0.0 else (math.nan if x.v == 0.0 else math.copysign(inf, x.v)))
    if _isc(y, 0.0):
        return C(math.nan)
    if _isc(x, 0.0):
        return ZERO
    if _isc(y, 1.0):
        return x
    if _isc(y, -1.0):
        return neg(x)
    if _eq(x, y):
        return ONE
    if x.op == "mul" and _eq(y, x.a):
        return x.b
    if x.op == "mul" and _eq(y, x.b):
        return x.a
    if _isc(x, 1.0):
        return unary("inv", y)
    if y.op == "inv":
        return mul(x, y.a)
    if y.op == "c" and x.op == "div" and x.b.op == "c" and y.v * x.b.v == 1.0:
        return x.a
    if y.op == "mul" and _eq(y.b, x):
        return unary("inv", y.a)
    if y.op == "mul" and _eq(y.a, x):
        return unary("inv", y.b)
    if x.op == "neg":
        return neg(div(x.a, y))
    if y.op == "neg":
        return neg(div(x, y.a))
    return E("div", x, y)


def power(x, y):
    if x.op == "c" and y.op == "c":
        return C(_safe(math.pow, x.v, y.v))
    if y.op == "c":
        if y.v == 0.0:
            return ONE
        if y.v == 1.0:
            return x
        if y.v == 2.0:
            return unary("sq", x)
        if y.v == 0.5:
            return unary("sqrt", x)
        if y.v == -1.0:
            return unary("inv", x)
        return E("cpow", x, y)
    return E("pow", x, y)


def unary(op, x):
    if op == "neg":
        return neg(x)
    if x.op == "c":
        return C(_safe(_UN[op], x.v))
    if op == "sq" and x.op == "neg":
        return unary("sq", x.a)
    if op == "inv" and x.op == "inv":
        return x.a
    return E(op, x)


def _topo(roots):
    """Nodes reachable from ``roots`` in dependency order (iterative: graphs are deep, e.g. products of 50 factors)."""
    order, seen, stack = [], set(), [(r, False) for r in reversed(roots)]
    while stack:
        n, done = stack.pop()
        if done:
            order.append(n)
            continue
        if id(n) in seen:
            continue
        seen.add(id(n))
        stack.append((n, True))
        if n.b is not None and n.op not in ("c", "s"):
            stack.append((n.b, False))
        if n.a is not None and n.op not in ("c", "s"):
            stack.append((n.a, False))
    return order


def _val(op, a, b=None):
    if op == "add":
        return a + b
    if op == "sub":
        return a - b
    if op == "mul":
        return a * b
    if op == "div":
        return a / b if b != 0.0 else (math.nan if a == 0.0 or a != a else math.copysign(inf, a) * math.copysign(1.0, b))
    if op in ("pow", "cpow"):
        return _safe(math.pow, a, b)
    return _safe(_UN[op], a)


def _partials(op, a, b, r):
    """(d r / d a, d r / d b) of one operation at numeric arguments."""
    if op == "add":
        return 1.0, 1.0
    if op == "sub":
        return 1.0, -1.0
    if op == "mul":
        return b, a
    if op == "div":
        return 1.0 / b, -r / b
    if op == "cpow":
        return b * _safe(math.pow, a, b - 1.0), 0.0
    if op == "pow":
        return b * _safe(math.pow, a, b - 1.0), (r * _safe(math.log, a))
    if op == "neg":
        return -1.0, None
    if op == "sq":
        return 2.0 * a, None
    if op == "sqrt":
        return _safe(lambda: 1.0 / (2.0 * r)), None
    if op == "sin":
        return math.cos(a), None
    if op == "cos":
        return -math.sin(a), None
    if op == "tan":
        return 1.0 + r * r, None
    if op == "exp":
        return r, None
    if op == "log":
        return _safe(lambda: 1.0 / a), None
    if op == "asin":
        return _safe(lambda: 1.0 / math.sqrt(1.0 - a * a)), None
    if op == "acos":
        return _safe(lambda: -1.0 / math.sqrt(1.0 - a * a)), None
    if op == "atan":
        return 1.0 / (1.0 + a * a), None
    if op == "sinh":
        return math.cosh(a), None
    if op == "cosh":
        return math.sinh(a), None
    if op == "tanh":
        return 1.0 - r * r, None
    if op == "fabs":
        return (1.0 if a > 0 else -1.0 if a < 0 else 0.0), None
    if op == "inv":
        return -r * r, None
    raise NotImplementedError(op)


def _partials2(op, a, b, r):
    """Second partials (r_aa, r_ab, r_bb) of one operation; ``None`` where the partial vanishes identically."""
    if op in ("add", "sub", "neg", "fabs"):
        return None, None, None
    if op == "mul":
        return None, 1.0, None
    if op == "div":
        return None, -1.0 / (b * b), 2.0 * r / (b * b)
    if op == "cpow":
        return b * (b - 1.0) * _safe(math.pow, a, b - 2.0), None, None
    if op == "sq":
        return 2.0, None, None
    if op == "sqrt":
        return _safe(lambda: -0.25 / (r * a)), None, None
    if op == "sin":
        return -math.sin(a), None, None
    if op == "cos":
        return -math.cos(a), None, None
    if op == "tan":
        return 2.0 * r * (1.0 + r * r), None, None
    if op == "exp":
        return r, None, None
    if op == "log":
        return _safe(lambda: -1.0 / (a * a)), None, None
    if op == "asin":
        return _safe(lambda: a / (1.0 - a * a) ** 1.5), None, None
    if op == "acos":
        return _safe(lambda: -a / (1.0 - a * a) ** 1.5), None, None
    if op == "atan":
        return -2.0 * a / (1.0 + a * a) ** 2, None, None
    if op == "sinh":
        return r, None, None
    if op == "cosh":
        return r, None, None
    if op == "tanh":
        return -2.0 * r * (1.0 - r * r), None, None
    if op == "inv":
        return 2.0 * r * r * r, None, None
    raise NotImplementedError(op)


def _symdiff(root, var):
    """Symbolic derivative d root / d var (forward mode over the graph, simplifying as it goes)."""
    d = {}
    for n in _topo([root]):
        if n.op == "c":
            d[id(n)] = ZERO
        elif n.op == "s":
            d[id(n)] = ONE if n is var else ZERO
        elif n.b is None:
            da, x = d[id(n.a)], n.a
            if _isc(da, 0.0):
                d[id(n)] = ZERO
                continue
            p = {
                "neg": lambda: C(-1.0), "sq": lambda: mul(C(2.0), x), "sqrt": lambda: div(ONE, mul(C(2.0), n)),
                "sin": lambda: unary("cos", x), "cos": lambda: neg(unary("sin", x)),
                "tan": lambda: add(ONE, unary("sq", n)), "exp": lambda: n, "log": lambda: unary("inv", x),
                "atan": lambda: unary("inv", add(ONE, unary("sq", x))), "sinh": lambda: unary("cosh", x),
                "cosh": lambda: unary("sinh", x), "tanh": lambda: sub(ONE, unary("sq", n)),
                "inv": lambda: neg(unary("sq", n)),
                "asin": lambda: unary("inv", unary("sqrt", sub(ONE, unary("sq", x)))),
                "acos": lambda: neg(unary("inv", unary("sqrt", sub(ONE, unary("sq", x))))),
            }[n.op]()
            d[id(n)] = mul(p, da)
        else:
            da, db = d[id(n.a)], d[id(n.b)]
            if n.op == "add":
                d[id(n)] = add(da, db)
            elif n.op == "sub":
                d[id(n)] = sub(da, db)
            elif n.op == "mul":
                d[id(n)] = add(mul(da, n.b), mul(n.a, db))
            elif n.op == "div":
                d[id(n)] = sub(div(da, n.b), mul(div(n, n.b), db))
            elif n.op == "cpow":
                d[id(n)] = mul(mul(n.b, power(n.a, C(n.b.v - 1.0))), da)
            else:
                raise NotImplementedError(n.op)
    return d[id(root)]


# --------------------------------------------------------------------------------------------------------------------
# matrices


def _elem(x):
    if isinstance(x, E):
        return x
    if isinstance(x, SX):
        if x.a.size != 1:
            raise ValueError("expected a scalar, got shape %s" % (x.a.shape,))
        return x.a.flat[0]
    return C(x)


def _obj(shape):
    return np.empty(shape, dtype=object)


class SX:
    __array_priority__ = 1e6
    __array_ufunc__ = None
    _numeric = False

    def __init__(self, *args):
        if len(args) == 0:
            self.a = _obj((0, 0))
        elif len(args) == 2 and all(isinstance(v, (int, np.integer)) for v in args):
            self.a = _obj(args)
            self.a[...] = ZERO
        elif len(args) == 1:
            self.a = _M(args[0]).a.copy()
        else:
            raise TypeError(args)

    # -- construction ------------------------------------------------------------------------------------------------
    @classmethod
    def _mk(cls, arr):
        m = cls.__new__(cls)
        m.a = arr
        return m

    @classmethod
    def sym(cls, name, *dims):
        if len(dims) == 3:
            return [cls.sym(f"{name}_{k}", dims[0], dims[1]) for k in range(dims[2])]
        r, c = (dims + (1, 1))[:2] if dims else (1, 1)
        if len(dims) == 1 and isinstance(dims[0], tuple):
            r, c = dims[0]
        a = _obj((r, c))
        for j in range(c):
            for i in range(r):
                a[i, j] = E("s", v=f"{name}_{i + j * r}" if r * c > 1 else name)
        return SX._mk(a)

    @classmethod
    def zeros(cls, *dims):
        if len(dims) == 1 and isinstance(dims[0], tuple):
            dims = dims[0]
        r, c = (tuple(dims) + (1,))[:2]
        a = _obj((r, c))
        a[...] = ZERO
        return cls._mk(a)

    @classmethod
    def ones(cls, *dims):
        m = cls.zeros(*dims)
        m.a[...] = ONE
        return m

    @classmethod
    def eye(cls, n):
        m = cls.zeros(n, n)
        for i in range(n):
            m.a[i, i] = ONE
        return m

    # -- shape -------------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.a.shape

    def size(self, *k):
        return self.a.shape if not k else self.a.shape[k[0] - 1]

    def size1(self):
        return self.a.shape[0]

    def size2(self):
        return self.a.shape[1]

    def numel(self):
        return self.a.size

    def __len__(self):
        return self.a.shape[0]

    @property
    def T(self):
        return _wrap(self.a.T.copy(), self)

    def reshape(self, shape):
        return _wrap(self.a.reshape(shape, order="F"), self)

    def is_constant(self):
        return all(e.op == "c" for e in self.a.flat)

    # -- indexing (CasADi: one index = linear, column-major; a slice returns a column) --------------------------------
    def __getitem__(self, k):
        if isinstance(k, tuple):
            i, j = k
            sub_ = self.a[_ix(i), :][:, _ix(j)]
            return _wrap(sub_.copy(), self)
        flat = self.a.reshape(-1, order="F")
        if isinstance(k, (int, np.integer)):
            out = _obj((1, 1))
            out[0, 0] = flat[k]
            return _wrap(out, self)
        sel = flat[k]
        return _wrap(sel.reshape(-1, 1).copy(), self)

    def __setitem__(self, k, v):
        v = _M(v)
        if isinstance(k, tuple):
            i, j = _ix(k[0]), _ix(k[1])
            rows = np.arange(self.a.shape[0])[i]
            cols = np.arange(self.a.shape[1])[j]
            src = v.a if v.a.size > 1 else np.broadcast_to(v.a, (len(rows), len(cols)))
            src = src.reshape(len(rows), len(cols)) if src.size == len(rows) * len(cols) else src
            for ii, r in enumerate(rows):
                for jj, c in enumerate(cols):
                    self.a[r, c] = src[ii, jj]
        else:
            n = self.a.size
            idx = [k % n] if isinstance(k, (int, np.integer)) else np.arange(n)[k]
            src = v.a.reshape(-1, order="F")
            for q_, q in enumerate(idx):
                self.a[q % self.a.shape[0], q // self.a.shape[0]] = src[q_ if src.size > 1 else 0]

    def __iter__(self):
        raise TypeError("SX is not iterable (as in CasADi)")

    # -- arithmetic ----------------------------------------------------------------------------------------------------
    def _bin(self, o, fn, swap=False):
        o = _M(o)
        x, y = (o, self) if swap else (self, o)
        xa, ya = x.a, y.a
        if xa.size == 1 and ya.size != 1:
            xa = np.broadcast_to(xa.reshape(1, 1), ya.shape)
        elif ya.size == 1 and xa.size != 1:
            ya = np.broadcast_to(ya.reshape(1, 1), xa.shape)
        elif xa.shape != ya.shape:
            raise ValueError(f"dimension mismatch {xa.shape} vs {ya.shape}")
        out = _obj(xa.shape)
        for idx in np.ndindex(*xa.shape):
            out[idx] = fn(xa[idx], ya[idx])
        return _wrap(out, x, y)

    def __add__(self, o): return self._bin(o, add)
    def __radd__(self, o): return self._bin(o, add, True)
    def __sub__(self, o): return self._bin(o, sub)
    def __rsub__(self, o): return self._bin(o, sub, True)
    def __mul__(self, o): return self._bin(o, mul)
    def __rmul__(self, o): return self._bin(o, mul, True)
    def __truediv__(self, o): return self._bin(o, div)
    def __rtruediv__(self, o): return self._bin(o, div, True)
    def __pow__(self, o): return self._bin(o, power)
    def __rpow__(self, o): return self._bin(o, power, True)
    def __matmul__(self, o): return mtimes(self, o)
    def __neg__(self): return self._un("neg")
    def __pos__(self): return self

    def _un(self, op):
        out = _obj(self.a.shape)
        for idx in np.ndindex(*self.a.shape):
            out[idx] = unary(op, self.a[idx])
        return _wrap(out, self)

    def sqrt(self): return self._un("sqrt")
    def sin(self): return self._un("sin")
    def cos(self): return self._un("cos")
    def tan(self): return self._un("tan")
    def exp(self): return self._un("exp")
    def log(self): return self._un("log")
    def arcsin(self): return self._un("asin")
    def arccos(self): return self._un("acos")
    def arctan(self): return self._un("atan")
    asin, acos, atan = arcsin, arccos, arctan
    def sinh(self): return self._un("sinh")
    def cosh(self): return self._un("cosh")
    def tanh(self): return self._un("tanh")
    def fabs(self): return self._un("fabs")
    __abs__ = fabs

    # -- numeric views (DM) ------------------------------------------------------------------------------------------
    def full(self):
        if not self.is_constant():
            raise TypeError("symbolic matrix has no numeric value")
        out = np.empty(self.a.shape)
        for idx in np.ndindex(*self.a.shape):
            out[idx] = self.a[idx].v
        return out

    toarray = full

    def __array__(self, dtype=None, copy=None):
        return self.full() if dtype is None else self.full().astype(dtype)

    def __float__(self):
        if self.a.size != 1:
            raise TypeError("only 1x1 matrices convert to float")
        return float(self.full().flat[0])

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        return bool(float(self))

    def __repr__(self):
        return f"{type(self).__name__}({self.a.tolist()!r})"

    def __eq__(self, o):  # numeric comparison for DM, identity otherwise (enough for the reference's use)
        if self.is_constant() and _M(o).is_constant():
            return np.array_equal(self.full(), _M(o).full()) if self.a.size != 1 else float(self) == float(_M(o))
        return self is o

    __hash__ = object.__hash__

    def __lt__(self, o): return float(self) < float(_M(o))
    def __le__(self, o): return float(self) <= float(_M(o))
    def __gt__(self, o): return float(self) > float(_M(o))
    def __ge__(self, o): return float(self) >= float(_M(o))


class DM(SX):
    """Numeric matrix: an ``SX`` whose entries are all constants."""
    _numeric = True

    def __init__(self, *args):
        SX.__init__(self, *args)
        if not self.is_constant():
            raise TypeError("DM from symbolic data")


def _ix(i):
    if isinstance(i, (int, np.integer)):
        return slice(i, i + 1) if i != -1 else slice(-1, None)
    if isinstance(i, slice):
        return i
    return np.asarray(i)


def _wrap(arr, *operands):
    cls = DM if all(getattr(o, "_numeric", True) for o in operands) and all(e.op == "c" for e in arr.flat) else SX
    return cls._mk(arr)


def _M(x):
    """Anything the reference hands to CasADi -> SX / DM (1-D numeric data becomes a column)."""
    if isinstance(x, SX):
        return x
    if isinstance(x, E):
        a = _obj((1, 1))
        a[0, 0] = x
        return SX._mk(a)
    if isinstance(x, (int, float, np.integer, np.floating, bool)):
        a = _obj((1, 1))
        a[0, 0] = C(x)
        return DM._mk(a)
    if isinstance(x, np.ndarray) and x.dtype != object:
        v = x.reshape(-1, 1) if x.ndim <= 1 else x
        a = _obj(v.shape)
        for idx in np.ndindex(*v.shape):
            a[idx] = C(v[idx])
        return DM._mk(a)
    if isinstance(x, (list, tuple, np.ndarray)):
        items = [_M(e) for e in x]
        if not items:
            return DM._mk(_obj((0, 1)))
        if all(m.a.size == 1 for m in items):
            a = _obj((len(items), 1))
            for i, m in enumerate(items):
                a[i, 0] = m.a.flat[0]
            return _wrap(a, *items)
        return vertcat(*[m.T for m in items])  # list of rows
    raise TypeError(f"cannot convert {type(x).__name__} to SX")


# --------------------------------------------------------------------------------------------------------------------
# free functions


def vertcat(*args):
    ms = [_M(a) for a in args]
    ms = [m for m in ms if m.a.size > 0]
    if not ms:
        return DM._mk(_obj((0, 1)))
    w = ms[0].a.shape[1]
    if any(m.a.shape[1] != w for m in ms):
        raise ValueError("vertcat: column counts differ: %s" % [m.a.shape for m in ms])
    return _wrap(np.concatenate([m.a for m in ms], axis=0), *ms)


def horzcat(*args):
    ms = [_M(a) for a in args]
    ms = [m for m in ms if m.a.size > 0]
    if not ms:
        return DM._mk(_obj((1, 0)))
    return _wrap(np.concatenate([m.a for m in ms], axis=1), *ms)


def mtimes(*args):
    if len(args) == 1:
        args = tuple(args[0])
    x, y = _M(args[0]), _M(args[1])
    for extra in args[2:]:
        return mtimes(mtimes(x, y), *args[2:])
    if x.a.size == 1 or y.a.size == 1:
        return x * y
    if x.a.shape[1] != y.a.shape[0]:
        raise ValueError(f"mtimes: {x.a.shape} x {y.a.shape}")
    n, m, K = x.a.shape[0], y.a.shape[1], x.a.shape[1]
    out = _obj((n, m))
    # CasADi's sparse product accumulates z(i,j) += x(i,k) y(k,j) with k ascending; exact-zero factors fold away
    nzx = [[k for k in range(K) if not _isc(x.a[i, k], 0.0)] for i in range(n)]
    for j in range(m):
        col = y.a[:, j]
        for i in range(n):
            acc = ZERO
            for k in nzx[i]:
                acc = add(acc, mul(x.a[i, k], col[k]))
            out[i, j] = acc
    return _wrap(out, x, y)


def kron(x, y):
    x, y = _M(x), _M(y)
    (p, q), (r, s) = x.a.shape, y.a.shape
    out = _obj((p * r, q * s))
    for i in range(p):
        for j in range(q):
            for k in range(r):
                for l in range(s):
                    out[i * r + k, j * s + l] = mul(x.a[i, j], y.a[k, l])
    return _wrap(out, x, y)


def diag(x):
    x = _M(x)
    if 1 in x.a.shape or x.a.size == 0:
        v = x.a.reshape(-1, order="F")
        out = _obj((v.size, v.size))
        out[...] = ZERO
        for i in range(v.size):
            out[i, i] = v[i]
        return _wrap(out, x)
    out = _obj((x.a.shape[0], 1))
    for i in range(x.a.shape[0]):
        out[i, 0] = x.a[i, i]
    return _wrap(out, x)


def solve(A, B, *_):
    """A \\ B.  Triangular systems by substitution (what CasADi does for them, exact for the reference's diagonal scaling
    matrices: entries 1/s); anything else numerically through LAPACK."""
    A, B = _M(A), _M(B)
    n = A.a.shape[0]
    if all(_isc(A.a[i, j], 0.0) for i in range(n) for j in range(n) if i != j):
        out = _obj(B.a.shape)
        for i in range(n):
            for j in range(B.a.shape[1]):
                out[i, j] = div(B.a[i, j], A.a[i, i])
        return _wrap(out, A, B)
    return DM(np.linalg.solve(A.full(), B.full()))


def sum1(x):
    x = _M(x)
    out = _obj((1, x.a.shape[1]))
    for j in range(x.a.shape[1]):
        acc = ZERO
        for i in range(x.a.shape[0]):
            acc = add(acc, x.a[i, j])
        out[0, j] = acc
    return _wrap(out, x)


def sum2(x):
    return sum1(_M(x).T).T


def sumsqr(x):
    acc = ZERO
    for e in _M(x).a.reshape(-1, order="F"):
        acc = add(acc, unary("sq", e))
    return _M(acc)


def dot(x, y):
    return sum1((_M(x) * _M(y)).reshape((-1, 1)))


def norm_2(x):
    return sumsqr(x).sqrt()


def _fn(name, np_name=None):
    def f(x):
        if isinstance(x, SX):
            return x._un(name)
        return getattr(np, np_name or name)(x)

    f.__name__ = name
    return f


sqrt, sin, cos, tan, exp, log = _fn("sqrt"), _fn("sin"), _fn("cos"), _fn("tan"), _fn("exp"), _fn("log")
asin, acos, atan = _fn("asin", "arcsin"), _fn("acos", "arccos"), _fn("atan", "arctan")
arcsin, arccos, arctan = asin, acos, atan
sinh, cosh, tanh, fabs = _fn("sinh"), _fn("cosh"), _fn("tanh"), _fn("fabs")


def transpose(x):
    return _M(x).T


def vec(x):
    return _M(x)[:]


def gradient(ex, wrt):
    """Symbolic gradient of a scalar expression; shaped like ``wrt``."""
    ex, wrt = _M(ex), _M(wrt)
    root = _elem(ex)
    out = _obj(wrt.a.shape)
    for idx in np.ndindex(*wrt.a.shape):
        out[idx] = _symdiff(root, wrt.a[idx])
    return SX._mk(out)


def jacobian(ex, wrt):
    ex, wrt = _M(ex)[:], _M(wrt)[:]
    out = _obj((ex.a.shape[0], wrt.a.shape[0]))
    for j in range(wrt.a.shape[0]):
        for i in range(ex.a.shape[0]):
            out[i, j] = _symdiff(ex.a[i, 0], wrt.a[j, 0])
    return SX._mk(out)


# --------------------------------------------------------------------------------------------------------------------
# Function: numeric evaluation of an expression graph


class Function:
    def __init__(self, name, ins, outs, names_in=None, names_out=None, *_):
        self.name = name
        self.ins = [_M(m) for m in ins]
        self.outs = [_M(m) for m in outs]
        self.names_in = list(names_in) if names_in and not isinstance(names_in, dict) else [f"i{k}" for k in range(len(ins))]
        self.names_out = list(names_out) if names_out and not isinstance(names_out, dict) else [f"o{k}" for k in range(len(outs))]
        roots = [e for m in self.outs for e in m.a.reshape(-1, order="F")]
        self.order = _topo(roots)
        self.slot = {id(n): i for i, n in enumerate(self.order)}
        self.in_slots = []
        for m in self.ins:
            sl = []
            for e in m.a.reshape(-1, order="F"):
                if e.op != "s":
                    raise ValueError("Function inputs must be purely symbolic")
                sl.append(self.slot.get(id(e), -1))
            self.in_slots.append(sl)
        known = {s for sl in self.in_slots for s in sl}
        for i, n in enumerate(self.order):
            if n.op == "s" and i not in known:
                raise ValueError(f"Function {name}: free variable {n.v}")
        self.prog = [(n.op, self.slot[id(n.a)] if n.op not in ("c", "s") else -1,
                      self.slot[id(n.b)] if (n.b is not None and n.op not in ("c", "s")) else -1, n.v) for n in self.order]
        self.out_slots = [[self.slot[id(e)] for e in m.a.reshape(-1, order="F")] for m in self.outs]

    def _numeric_inputs(self, args):
        vals = []
        for m, a in zip(self.ins, args):
            v = _M(a).full().reshape(-1, order="F") if not isinstance(a, np.ndarray) else np.asarray(a, float).reshape(-1, order="F")
            if v.size == 1 and m.a.size > 1:
                v = np.full(m.a.size, v[0])
            if v.size != m.a.size:
                raise ValueError(f"Function {self.name}: input of {v.size} elements for {m.a.shape}")
            vals.append(v)
        return vals

    def _run(self, vals):
        w = [0.0] * len(self.prog)
        for sl, v in zip(self.in_slots, vals):
            for s, x in zip(sl, v):
                if s >= 0:
                    w[s] = float(x)
        for i, (op, a, b, v) in enumerate(self.prog):
            if op == "c":
                w[i] = v
            elif op == "s":
                pass
            elif b >= 0:
                w[i] = _val(op, w[a], w[b])
            else:
                w[i] = _val(op, w[a])
        return w

    def __call__(self, *args, **kw):
        if kw:
            args = [kw[n] for n in self.names_in]
        w = self._run(self._numeric_inputs(args))
        res = []
        for m, sl in zip(self.outs, self.out_slots):
            res.append(DM(np.array([w[s] for s in sl], float).reshape(m.a.shape, order="F")))
        if kw:
            return dict(zip(self.names_out, res))
        return res[0] if len(res) == 1 else tuple(res)

    # -- helpers beyond CasADi's API, used by oracle/refrun/run_reference.py -------------------------------------------
    def forward_sparse(self, args, wrt=0):
        """Values of all outputs and, per output element, {column of input ``wrt`` -> derivative}: forward-mode AD with
        sparse tangents.  A column is present iff the output depends on it structurally (every operation passes on the
        union of its arguments' dependencies, which is how CasADi propagates Jacobian sparsity)."""
        w = self._run(self._numeric_inputs(args))
        d = [None] * len(self.prog)
        for col, s in enumerate(self.in_slots[wrt]):
            if s >= 0:
                d[s] = {col: 1.0}
        for i, (op, a, b, v) in enumerate(self.prog):
            if op in ("c", "s"):
                continue
            da = d[a]
            db = d[b] if b >= 0 else None
            if da is None and db is None:
                continue
            pa, pb = _partials(op, w[a], w[b] if b >= 0 else None, w[i])
            t = {}
            if da is not None:
                for c_, x in da.items():
                    t[c_] = pa * x
            if db is not None:
                for c_, x in db.items():
                    t[c_] = t[c_] + pb * x if c_ in t else pb * x
            d[i] = t
        outs = []
        for m, sl in zip(self.outs, self.out_slots):
            outs.append((np.array([w[s] for s in sl], float), [d[s] or {} for s in sl]))
        return outs

    def forward2_sparse(self, args, wrt=0):
        """Like ``forward_sparse`` plus, per output element, the lower triangle of its Hessian in input ``wrt`` as
        {(row, col): value} with row >= col (second-order forward mode with sparse tangents; an entry is present iff it
        is structurally non-zero: products of the arguments' dependencies wherever the operation has a second partial)."""
        w = self._run(self._numeric_inputs(args))
        n = len(self.prog)
        d, h = [None] * n, [None] * n
        for col, s in enumerate(self.in_slots[wrt]):
            if s >= 0:
                d[s] = {col: 1.0}
        for i, (op, a, b, v) in enumerate(self.prog):
            if op in ("c", "s"):
                continue
            da = d[a]
            db = d[b] if b >= 0 else None
            if da is None and db is None:
                continue
            wa, wb = w[a], (w[b] if b >= 0 else None)
            pa, pb = _partials(op, wa, wb, w[i])
            t, H = {}, {}
            if da is not None:
                for c_, x in da.items():
                    t[c_] = pa * x
                if h[a]:
                    for k_, x in h[a].items():
                        H[k_] = pa * x
            if db is not None:
                for c_, x in db.items():
                    t[c_] = t[c_] + pb * x if c_ in t else pb * x
                if h[b]:
                    for k_, x in h[b].items():
                        H[k_] = H[k_] + pb * x if k_ in H else pb * x
            raa, rab, rbb = _partials2(op, wa, wb, w[i])

            def outer(c, u, v_, sym):
                for p_, x in u.items():
                    for q_, y in v_.items():
                        val = c * x * y
                        if sym and p_ == q_:
                            val = 2.0 * val
                        k_ = (p_, q_) if p_ >= q_ else (q_, p_)
                        H[k_] = H[k_] + val if k_ in H else val

            if raa is not None and da is not None:
                # full outer product folded onto the lower triangle: off-diagonal pairs appear twice in the loop
                for p_, x in da.items():
                    for q_, y in da.items():
                        if p_ >= q_:
                            val = raa * x * y
                            H[(p_, q_)] = H[(p_, q_)] + val if (p_, q_) in H else val
            if rbb is not None and db is not None:
                for p_, x in db.items():
                    for q_, y in db.items():
                        if p_ >= q_:
                            val = rbb * x * y
                            H[(p_, q_)] = H[(p_, q_)] + val if (p_, q_) in H else val
            if rab is not None and da is not None and db is not None:
                outer(rab, da, db, True)
            d[i], h[i] = t, (H or None)
        outs = []
        for m, sl in zip(self.outs, self.out_slots):
            outs.append((np.array([w[s] for s in sl], float), [d[s] or {} for s in sl], [h[s] or {} for s in sl]))
        return outs



def integrator(name, plugin, dae, opts=None, *rest):
    """Stand-in for ``ca.integrator(name, "idas", {"x", "t", "ode"}, {"t0", "tf"})`` (mpopt.py:3869-3877): the reference
    only integrates polynomials in ``t`` with it, so the quadrature is done EXACTLY (96-point Gauss-Legendre, exact to
    degree 191) instead of with IDAS at its default tolerances."""
    opts = opts or {}
    t0, tf = float(opts.get("t0", 0.0)), float(opts.get("tf", 1.0))
    ode = Function(name + "_ode", [dae["t"]], [dae["ode"]])  # raises if the right-hand side depends on x
    xs, ws = np.polynomial.legendre.leggauss(96)

    def run(x0=0.0, **_):
        mid, half = 0.5 * (t0 + tf), 0.5 * (tf - t0)
        s = math.fsum(w * float(ode(mid + half * x)) for x, w in zip(xs, ws)) * half
        return {"xf": DM(float(_M(x0)) + s)}

    return run


class _NlpSolver:
    """``ca.nlpsol(name, "ipopt", {x, f, g, p}, opts)``: IPOPT is not installable either, so a call hands the NLP's
    evaluators (values and sparse first / second derivatives of the recorded graph) to the interior-point method of
    ``mpopt_b200.ipm`` -- the five callbacks IPOPT would get -- and returns CasADi's result dictionary."""

    def __init__(self, name, plugin, nlp, opts):
        self.name, self.plugin, self.nlp, self.opts = name, plugin, nlp, dict(opts or {})
        x = _M(nlp["x"])
        self.has_p = "p" in nlp and _M(nlp["p"]).a.size > 0
        g = _M(nlp["g"]) if "g" in nlp else DM._mk(_obj((0, 1)))
        ins = [x, _M(nlp["p"])] if self.has_p else [x]
        self.fn = Function(name, ins, [g, _M(nlp["f"])])
        self.n, self.m = x.a.size, g.a.size
        self.stats_ = {}

    def stats(self):
        return self.stats_

    def __call__(self, x0=None, p=None, lbx=-inf, ubx=inf, lbg=-inf, ubg=inf, lam_x0=None, lam_g0=None, **_):
        import scipy.sparse as sp

        from mpopt_b200.ipm import solve_nlp

        n, m = self.n, self.m
        vec = lambda v, k: np.broadcast_to(np.asarray(_M(v).full() if isinstance(v, SX) else v, float).ravel(), (k,)).copy() \
            if np.size(v) in (1, k) else np.asarray(v, float).ravel()
        pv = vec(p, _M(self.nlp["p"]).a.size) if self.has_p else None
        cache = {}

        def at(xx):
            key = xx.tobytes()
            if cache.get("key") != key:
                (g, dg, hg), (f, df, hf) = self.fn.forward2_sparse([xx, pv] if self.has_p else [xx], wrt=0)
                cache.update(key=key, g=g, dg=dg, hg=hg, f=float(f[0]), df=df[0], hf=hf[0])
            return cache

        def grad_f(xx):
            out = np.zeros(n)
            for c_, v in at(xx)["df"].items():
                out[c_] = v
            return out

        def jac_g(xx):
            rows, cols, vals = [], [], []
            for i, row in enumerate(at(xx)["dg"]):
                for c_, v in row.items():
                    rows.append(i), cols.append(c_), vals.append(v)
            return sp.csr_matrix((vals, (rows, cols)), shape=(m, n))

        def hess_l(xx, lam_f, lam_g):
            c = at(xx)
            H = {k_: lam_f * v for k_, v in c["hf"].items()}
            for l, h in zip(lam_g, c["hg"]):
                for k_, v in h.items():
                    H[k_] = H.get(k_, 0.0) + l * v
            ks = list(H)
            return sp.csr_matrix(([H[k_] for k_ in ks], ([k_[0] for k_ in ks], [k_[1] for k_ in ks])), shape=(n, n))

        tol = float(self.opts.get("ipopt.tol", 1e-8))
        r = solve_nlp(lambda xx: at(xx)["f"], grad_f, lambda xx: at(xx)["g"].copy(), jac_g, hess_l, vec(x0, n),
                      vec(lbx, n), vec(ubx, n), vec(lbg, m), vec(ubg, m), tol=tol,
                      max_iter=int(self.opts.get("ipopt.max_iter", 3000)),
                      acceptable_tol=float(self.opts.get("ipopt.acceptable_tol", 1e-6)))
        self.stats_ = {"success": bool(r.success), "iter_count": int(r.iter),
                       "return_status": "Solve_Succeeded" if r.success else "Maximum_Iterations_Exceeded"}
        return {"x": DM(r.x), "f": DM(r.f), "g": DM(r.g), "lam_g": DM(r.lam_g), "lam_x": DM(r.lam_x),
                "lam_p": DM(np.zeros(0 if pv is None else pv.size))}


def nlpsol(name, plugin, nlp, opts=None):
    return _NlpSolver(name, plugin, nlp, opts)
