"""Shared helpers for the parity tests."""
import numpy as np

RTOL = 1e-10  # north_star: float64 residual / Jacobian values within 1e-10 relative


def assert_close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert not np.isnan(a).any(), f"{what}: NaN in result"
    assert err.size == 0 or err.max() <= rtol, f"{what}: max rel err {err.max():.3e} at {int(err.argmax())}"


def random_point(oracle, seed=20261017, tf=None, dirichlet=False):
    """Seeded evaluation point in the layout of z, plus segment widths (SURVEY.md 8d 'Input values')."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, oracle.n_z)
    for ph in range(oracle.P):
        z[oracle.colT0(ph)] = 0.25 * ph + (0.1 if ph else 0.0)
        z[oracle.colTF(ph)] = (tf if tf is not None else 2.0) + 1.5 * ph
        for m in range(oracle.na):
            z[oracle.colA(ph, m)] = rng.uniform(0.2, 1.2)
    if dirichlet:
        p = np.concatenate([rng.dirichlet(np.ones(oracle.K)) for _ in range(oracle.P)])
    else:
        p = oracle.seg_width_params()
    return z, p


def eval_expr(outputs, env):
    """Numerically evaluate traced expressions (mpopt_b200.trace.Expr) at a point: test-only interpreter."""
    import math

    from mpopt_b200 import trace as tr

    val = {}
    fun = {"neg": lambda a: -a, "sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos,
           "tan": math.tan, "asin": math.asin, "acos": math.acos, "atan": math.atan, "sinh": math.sinh,
           "cosh": math.cosh, "tanh": math.tanh, "fabs": abs, "sign": lambda a: (a > 0) - (a < 0), "sq": lambda a: a * a}
    for e in tr.topo([tr.as_expr(o) for o in outputs]):
        if e.op == "const":
            val[e.id] = e.value
        elif e.op == "var":
            val[e.id] = env[e.name]
        elif e.op in ("add", "sub", "mul", "div", "pow"):
            a, b = (val[x.id] for x in e.args)
            val[e.id] = {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b if e.op == "div" else 0.0,
                         "pow": a ** b if e.op == "pow" else 0.0}[e.op]
        else:
            val[e.id] = fun[e.op](val[e.args[0].id])
    return [val[tr.as_expr(o).id] for o in outputs]
