"""Trace user Python callables into an expression DAG, differentiate it, emit CUDA.

Replaces what CasADi's SX layer does for the reference when
``get_discretized_dynamics_constraints_and_cost_matrices`` calls the user's
``dynamics / path_constraints / running_costs`` with symbolic arguments
(/root/reference/mpopt/mpopt.py:186-206) and ``ca.nlpsol`` differentiates the
result (:757).  Here each callable is traced ONCE per phase (not once per node):
the node-local function ``(x, u, t, a) -> f`` is recorded as a hash-consed DAG,
differentiated symbolically in reverse mode, and emitted as straight-line
``double`` CUDA code that the hand-written kernels in ``csrc/`` inline per node.

Construction-time folds match the SX ones that change the *structural* Jacobian
pattern: ``0*x -> 0``, ``x+0 -> x``, ``1*x -> x``,
``x-x -> 0``, ``x/x -> 1``, ``x**0 -> 1``, ``x**1 -> x`` and constant folding.
"""
from __future__ import annotations

import math
import numbers

import numpy as np

_UNARY = ("neg", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan",
          "sinh", "cosh", "tanh", "fabs", "sign", "sq")
_BINARY = ("add", "sub", "mul", "div", "pow")

_FOLD1 = {
    "neg": np.negative, "sqrt": np.sqrt, "exp": np.exp, "log": np.log, "sin": np.sin,
    "cos": np.cos, "tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
    "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "fabs": np.abs,
    "sign": np.sign, "sq": np.square,
}


def _fold(fn, *vals):
    """Constant folding in IEEE arithmetic: out-of-domain arguments, overflow and division by zero give nan / inf
    like CasADi's SX folds do (the Python math module would raise instead and abort the trace)."""
    with np.errstate(all="ignore"):
        return float(fn(*(np.float64(v) for v in vals)))


class Expr:
    """One DAG node.  ``op`` is 'const', 'var', or an operator name; nodes are interned."""

    __slots__ = ("op", "args", "value", "name", "id")
    __array_priority__ = 3000.0
    _intern: dict = {}
    _count = 0

    def __new__(cls, op, args=(), value=None, name=None):
        key = (op, tuple(a.id for a in args), value, name)
        hit = cls._intern.get(key)
        if hit is not None:
            return hit
        self = object.__new__(cls)
        self.op, self.args, self.value, self.name = op, tuple(args), value, name
        self.id = cls._count
        Expr._count += 1
        cls._intern[key] = self
        return self

    @classmethod
    def reset_interning(cls):
        """Forget every interned node except the shared constants (call between traced problems in a long-lived
        process: the table otherwise grows with every OCP).  Expressions created before the reset stay valid but no
        longer unify with new ones, so never mix the two in one program."""
        keep = {k: v for k, v in cls._intern.items() if v.op == "const" and v.value in (0.0, 1.0)}
        cls._intern.clear()
        cls._intern.update(keep)

    # ---- predicates
    @property
    def is_const(self):
        return self.op == "const"

    def is_value(self, v):
        return self.op == "const" and self.value == v

    def __repr__(self):
        if self.op == "const":
            return repr(self.value)
        if self.op == "var":
            return self.name
        return f"{self.op}({', '.join(map(repr, self.args))})"

    def __hash__(self):
        return self.id

    def __bool__(self):
        raise TypeError("traced expressions have no truth value (data-dependent branching is not traceable)")

    # ---- numpy interop
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        n = ufunc.__name__
        two = {"add": add, "subtract": sub, "multiply": mul, "true_divide": div, "divide": div, "power": power}
        if n in two:
            return two[n](*inputs)
        one = {"negative": "neg", "absolute": "fabs", "square": "sq", "arccos": "acos", "arcsin": "asin",
               "arctan": "atan"}
        n = one.get(n, n)
        if n == "positive":
            return self
        if len(inputs) == 1 and n in _UNARY:
            return unary(n, self)
        return NotImplemented

    # ---- operators
    # a list operand (ca.Vec, what vertcat / slicing return) takes over: scalar (op) vector is element-wise
    def __add__(self, o): return NotImplemented if isinstance(o, list) else add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return NotImplemented if isinstance(o, list) else sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return NotImplemented if isinstance(o, list) else mul(self, o)
    def __rmul__(self, o): return mul(o, self)
    def __truediv__(self, o): return NotImplemented if isinstance(o, list) else div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __pow__(self, o): return power(self, o)
    def __rpow__(self, o): return power(o, self)
    def __neg__(self): return unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return unary("fabs", self)

    # ---- elementary functions (found by the ``ca`` shim and by numpy ufuncs through duck typing)
    def sqrt(self): return unary("sqrt", self)
    def exp(self): return unary("exp", self)
    def log(self): return unary("log", self)
    def sin(self): return unary("sin", self)
    def cos(self): return unary("cos", self)
    def tan(self): return unary("tan", self)
    def asin(self): return unary("asin", self)
    def acos(self): return unary("acos", self)
    def atan(self): return unary("atan", self)
    def sinh(self): return unary("sinh", self)
    def cosh(self): return unary("cosh", self)
    def tanh(self): return unary("tanh", self)
    def fabs(self): return unary("fabs", self)
    arccos, arcsin, arctan = acos, asin, atan


def const(v) -> Expr:
    v = float(v)
    if v == 0.0:
        v = 0.0  # -0.0 and 0.0 intern to the same node
    return Expr("const", (), v)


def var(name: str) -> Expr:
    return Expr("var", (), None, name)


ZERO = const(0.0)
ONE = const(1.0)


def as_expr(v) -> Expr:
    if isinstance(v, Expr):
        return v
    if isinstance(v, np.ndarray) and v.dtype != object and v.size == 1:
        v = v.reshape(-1)[0]
    if isinstance(v, (numbers.Real, np.floating, np.integer)):
        return const(v)
    raise TypeError(f"cannot trace a value of type {type(v)!r}")


def unary(op: str, a) -> Expr:
    a = as_expr(a)
    if a.is_const:
        return const(_fold(_FOLD1[op], a.value))
    if op == "neg" and a.op == "neg":
        return a.args[0]
    return Expr(op, (a,))


def add(a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if a.is_const and b.is_const:
        return const(a.value + b.value)
    if a.is_value(0.0):
        return b
    if b.is_value(0.0):
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    return Expr("add", (a, b))


def sub(a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if a.is_const and b.is_const:
        return const(a.value - b.value)
    if a is b:
        return ZERO
    if b.is_value(0.0):
        return a
    if a.is_value(0.0):
        return unary("neg", b)
    if b.op == "neg":
        return add(a, b.args[0])
    return Expr("sub", (a, b))


def mul(a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if a.is_const and b.is_const:
        return const(_fold(np.multiply, a.value, b.value))
    if a.is_value(0.0) or b.is_value(0.0):
        return ZERO
    if a.is_value(1.0):
        return b
    if b.is_value(1.0):
        return a
    if a.is_value(-1.0):
        return unary("neg", b)
    if b.is_value(-1.0):
        return unary("neg", a)
    if a is b:
        return Expr("sq", (a,))
    return Expr("mul", (a, b))


def div(a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if a.is_const and b.is_const:
        return const(_fold(np.divide, a.value, b.value))
    if a.is_value(0.0):
        return ZERO
    if b.is_value(1.0):
        return a
    if a is b:
        return ONE
    return Expr("div", (a, b))


def power(a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if a.is_const and b.is_const:
        return const(_fold(np.power, a.value, b.value))
    if b.is_const:
        e = b.value
        if e == 0.0:
            return ONE
        if e == 1.0:
            return a
        if e == 2.0:
            return Expr("sq", (a,))
        if e == 0.5:
            return unary("sqrt", a)
    return Expr("pow", (a, b))


# --------------------------------------------------------------------------- differentiation
def _partials(e: Expr):
    """d e / d arg for every argument of ``e`` (as Exprs)."""
    op, a = e.op, e.args
    if op == "add":
        return (ONE, ONE)
    if op == "sub":
        return (ONE, const(-1.0))
    if op == "mul":
        return (a[1], a[0])
    if op == "div":
        return (div(ONE, a[1]), unary("neg", div(e, a[1])))
    if op == "pow":
        x, y = a
        dx = mul(y, power(x, sub(y, ONE)))
        dy = ZERO if y.is_const else mul(e, unary("log", x))
        return (dx, dy)
    x = a[0]
    if op == "neg":
        return (const(-1.0),)
    if op == "sq":
        return (mul(const(2.0), x),)
    if op == "sqrt":
        return (div(const(0.5), e),)
    if op == "exp":
        return (e,)
    if op == "log":
        return (div(ONE, x),)
    if op == "sin":
        return (unary("cos", x),)
    if op == "cos":
        return (unary("neg", unary("sin", x)),)
    if op == "tan":
        return (add(ONE, mul(e, e)),)
    if op == "asin":
        return (div(ONE, unary("sqrt", sub(ONE, mul(x, x)))),)
    if op == "acos":
        return (unary("neg", div(ONE, unary("sqrt", sub(ONE, mul(x, x))))),)
    if op == "atan":
        return (div(ONE, add(ONE, mul(x, x))),)
    if op == "sinh":
        return (unary("cosh", x),)
    if op == "cosh":
        return (unary("sinh", x),)
    if op == "tanh":
        return (sub(ONE, mul(e, e)),)
    if op == "fabs":
        return (unary("sign", x),)
    if op == "sign":
        return (ZERO,)
    raise NotImplementedError(op)


def topo(outputs):
    """Nodes reachable from ``outputs`` in dependency order: iterative DFS post-order, so the order depends
    only on the DAG's structure (not on creation ids shared with earlier traces in this process)."""
    order, done = [], set()
    for root in outputs:
        if root.id in done:
            continue
        stack = [(root, 0)]
        while stack:
            e, i = stack.pop()
            if e.id in done:
                continue
            if i < len(e.args):
                stack.append((e, i + 1))
                if e.args[i].id not in done:
                    stack.append((e.args[i], 0))
            else:
                done.add(e.id)
                order.append(e)
    return order


def gradient(out: Expr, wrt):
    """Reverse-mode symbolic gradient of scalar ``out`` w.r.t. the variables in ``wrt`` -> list of Expr."""
    out = as_expr(out)
    adj = {out.id: ONE}
    for e in reversed(topo([out])):
        bar = adj.get(e.id)
        if bar is None or e.op in ("const", "var"):
            continue
        for arg, p in zip(e.args, _partials(e)):
            if arg.is_const:
                continue
            c = mul(bar, p)
            if c.is_value(0.0):
                continue
            adj[arg.id] = add(adj[arg.id], c) if arg.id in adj else c
    return [adj.get(v.id, ZERO) for v in wrt]


def depends_on(e: Expr, v: Expr) -> bool:
    return any(n is v for n in topo([e]))


# --------------------------------------------------------------------------- code generation
_C_FUN = {"sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin", "cos": "cos", "tan": "tan", "asin": "asin",
          "acos": "acos", "atan": "atan", "sinh": "sinh", "cosh": "cosh", "tanh": "tanh", "fabs": "fabs"}


def c_literal(v: float) -> str:
    if math.isinf(v):
        return "(1.0/0.0)" if v > 0 else "(-1.0/0.0)"
    if math.isnan(v):
        return "(0.0/0.0)"
    s = repr(float(v))
    if "e" not in s and "." not in s:
        s += ".0"
    return s


def emit_c(outputs, var_ref, indent="  ", prefix="v"):
    """Straight-line C for a list of output Exprs.

    ``var_ref`` maps variable name -> C expression.  Returns (lines, refs) where ``refs[i]`` is the C
    expression holding ``outputs[i]``.
    """
    outputs = [as_expr(o) for o in outputs]
    name = {}
    lines = []
    for e in topo(outputs):
        if e.op == "const":
            name[e.id] = c_literal(e.value)
            continue
        if e.op == "var":
            name[e.id] = var_ref[e.name]
            continue
        a = [name[x.id] for x in e.args]
        if e.op == "add":
            rhs = f"{a[0]} + {a[1]}"
        elif e.op == "sub":
            rhs = f"{a[0]} - {a[1]}"
        elif e.op == "mul":
            rhs = f"{a[0]} * {a[1]}"
        elif e.op == "div":
            rhs = f"{a[0]} / {a[1]}"
        elif e.op == "neg":
            rhs = f"-{a[0]}"
        elif e.op == "sq":
            rhs = f"{a[0]} * {a[0]}"
        elif e.op == "sign":
            rhs = f"(double)(({a[0]} > 0.0) - ({a[0]} < 0.0))"
        elif e.op == "pow":
            y = e.args[1]
            if y.is_const and float(y.value).is_integer() and abs(y.value) <= 8:
                n = int(abs(y.value))
                prod = " * ".join([a[0]] * n)
                rhs = prod if y.value > 0 else f"1.0 / ({prod})"
            else:
                rhs = f"pow({a[0]}, {a[1]})"
        else:
            rhs = f"{_C_FUN[e.op]}({a[0]})"
        nm = f"{prefix}{len(lines)}"
        lines.append(f"{indent}const double {nm} = {rhs};")
        name[e.id] = nm
    return lines, [name[o.id] for o in outputs]


def flatten(out):
    """Flatten a user callable's return value (scalar / list / nested / ndarray) -> list, or None."""
    if out is None:
        return None
    if isinstance(out, Expr) or isinstance(out, (numbers.Real, np.floating, np.integer)):
        return [out]
    if isinstance(out, np.ndarray):
        out = out.ravel().tolist()
    res = []
    for o in out:
        res.extend(flatten(o))
    return res
