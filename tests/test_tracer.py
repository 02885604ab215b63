"""Host logic: the product tracer (DAG + symbolic reverse-mode AD + CUDA codegen) against the oracle's
independent dual numbers, plus the SX-style folding rules that decide the structural pattern."""
import numpy as np
import pytest

from helpers import eval_expr
from mpopt_b200 import ca, trace as tr
from mpopt_b200.problems import EXAMPLES, REGISTRY

ALL_PROBLEMS = {**REGISTRY, **EXAMPLES}
from mpopt_b200.program import Program
from oracle.dual import Dual, Vec, flatten


def test_folding_rules():
    x, y = tr.var("fx"), tr.var("fy")
    assert (0 * x).is_value(0.0) and (x * 0.0).is_value(0.0)
    assert (x + 0) is x and (0 + x) is x and (x - 0) is x and (1 * x) is x and (x / 1) is x
    assert (x - x).is_value(0.0) and (x / x).is_value(1.0)
    assert (x ** 0).is_value(1.0) and (x ** 1) is x
    assert (x * y) is (x * y)  # hash-consing
    assert -(-x) is x
    assert tr.as_expr(2.0) * 3.0 is tr.const(6.0)
    g = tr.gradient(x * y + ca.sin(x), [x, y])
    assert eval_expr(g, {"fx": 0.3, "fy": 2.0}) == pytest.approx([2.0 + np.cos(0.3), 0.3])
    assert tr.gradient(x * x, [y])[0].is_value(0.0)


def test_numpy_and_shim_interop():
    x = tr.var("nx0")
    e = np.sqrt(x) + np.float64(2.0) * x + ca.exp(x) + np.cos(x) ** 2 + abs(x)
    v = eval_expr([e], {"nx0": 0.7})[0]
    assert v == pytest.approx(np.sqrt(0.7) + 1.4 + np.exp(0.7) + np.cos(0.7) ** 2 + 0.7)
    vec = ca.vertcat(x, 2 * x, [x * x])
    assert len(vec) == 3 and len(vec[:2]) == 2
    s = 3.0 * vec[:2]
    assert eval_expr(list(s), {"nx0": 2.0}) == [6.0, 12.0]
    with pytest.raises(TypeError):
        bool(x > 0) if hasattr(x, "__gt__") else bool(x)


@pytest.mark.parametrize("name", sorted(ALL_PROBLEMS))
def test_partials_and_pattern_match_oracle_duals(name):
    ocp = ALL_PROBLEMS[name]()
    prog = Program(ocp)
    rng = np.random.default_rng(3)
    nx, nu, na = ocp.nx, ocp.nu, ocp.na
    for ph, pp in enumerate(prog.phases):
        xv, uv, av, tv = rng.uniform(0.6, 1.4, nx), rng.uniform(0.6, 1.4, nu), rng.uniform(0.6, 1.4, na), 0.37
        env = {v.name: xv[i] for i, v in enumerate(pp.x)}
        env.update({v.name: uv[i] for i, v in enumerate(pp.u)})
        env.update({v.name: av[i] for i, v in enumerate(pp.a)})
        env[pp.t.name] = tv
        one = np.ones(1)
        x = Vec(Dual([xv[s]], {s: one}) for s in range(nx))
        u = Vec(Dual([uv[c]], {nx + c: one}) for c in range(nu))
        a = Vec(Dual([av[m]], {nx + nu + m: one}) for m in range(na))
        t = Dual([tv], {"t": one})
        for outs, entries, dts, fn in ((pp.f, pp.jf, pp.ft, ocp.get_dynamics(ph)),
                                       (pp.c, pp.jc, pp.ct, ocp.get_path_constraints(ph))):
            if not outs:
                continue
            ref = flatten(fn(x, u, t, a))
            vals = eval_expr(outs, env)
            pattern = {(r, v) for r, v, _ in entries}
            for r, d in enumerate(ref):
                if isinstance(d, Dual):
                    assert vals[r] == pytest.approx(float(d.val[0]), rel=1e-13, abs=1e-13)
                    assert {(r, k) for k in d.der if k != "t"} == {e for e in pattern if e[0] == r}
                    assert ("t" in d.der) == (not dts[r].is_value(0.0))
                    if "t" in d.der:
                        assert eval_expr([dts[r]], env)[0] == pytest.approx(float(d.der["t"][0]), rel=1e-12)
                else:
                    assert vals[r] == pytest.approx(float(d)) and not {e for e in pattern if e[0] == r}
            for r, v, e in entries:
                assert eval_expr([e], env)[0] == pytest.approx(float(ref[r].der[v][0]), rel=1e-12, abs=1e-14)


def test_program_source_is_deterministic_and_order_independent():
    keys1 = {n: Program(f()).key() for n, f in REGISTRY.items()}
    keys2 = {n: Program(f()).key() for n, f in reversed(list(REGISTRY.items()))}
    assert keys1 == keys2
    assert len(set(keys1.values())) == len(keys1)
    src = Program(REGISTRY["moon_lander"]()).cuda_source()
    assert "struct MPX_PHASE_NAME(0)" in src and "__device__" in src and "jf[1] = 1.0;" in src


def test_row_layouts_moon_lander():
    pp = Program(REGISTRY["moon_lander"]()).phases[0]
    assert pp.pat_f() == [[0, 1, 0], [0, 0, 1]] and pp.f_nz == [True, True]
    assert pp.f_row_layout(0) == ([], [("x", 1), ("T0", 0), ("TF", 0)])
    assert pp.f_row_layout(1) == ([], [("u", 0), ("T0", 0), ("TF", 0)])
    assert pp.tc_row_layout(0) == [0] and pp.tc_row_layout(1) == [1]


def test_constant_folding_follows_ieee_like_sx():
    """Out-of-domain constants fold to nan / inf instead of aborting the trace (CasADi's SX folds the same way)."""
    import math

    from mpopt_b200 import trace as tr

    assert math.isnan(tr.unary("sqrt", -1.0).value) and math.isnan(tr.unary("acos", 2.0).value)
    assert tr.unary("log", 0.0).value == -math.inf and tr.unary("exp", 1000.0).value == math.inf
    assert tr.div(1.0, 0.0).value == math.inf and tr.div(-1.0, 0.0).value == -math.inf
    assert tr.power(0.0, -1.0).value == math.inf and math.isnan(tr.div(0.0, 0.0).value)
    lines, refs = tr.emit_c([tr.mul(tr.var("x"), tr.unary("sqrt", -1.0))], {"x": "x[0]"})
    assert "(0.0/0.0)" in "".join(lines)


def test_interning_is_scoped_to_a_program():
    """The intern table does not grow with every traced OCP, and the generated source does not depend on what was
    traced before (the hash selects the AOT object)."""
    from mpopt_b200 import problems, trace as tr
    from mpopt_b200.program import Program

    k1 = Program(problems.moon_lander()).key()
    n1 = len(tr.Expr._intern)
    Program(problems.van_der_pol())
    Program(problems.hyper_sensitive())
    assert len(tr.Expr._intern) == n1 <= 4
    assert Program(problems.moon_lander()).key() == k1
