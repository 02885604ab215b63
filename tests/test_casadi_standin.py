"""The CasADi stand-in of oracle/refrun (test infrastructure behind the reference-generated fixtures) on its own:
indexing conventions, SX's construction-time simplifications, numeric evaluation, sparse forward AD against finite
differences, the exact ``integrator``."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ca():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refrun", "stubs"))
    try:
        import casadi

        if not casadi.__file__.startswith(os.path.join(ROOT, "oracle")):
            pytest.skip("a real casadi is installed")
        yield casadi
    finally:
        sys.path.remove(os.path.join(ROOT, "oracle", "refrun", "stubs"))


def test_indexing_is_column_major_like_casadi(ca):
    X = ca.SX.sym("x", 3, 2)
    assert X[:].shape == (6, 1) and X[:].a[3, 0] is X.a[0, 1]  # vec() stacks columns
    assert X[4].a[0, 0] is X.a[1, 1]                           # one index = linear, column-major
    assert X[1, :].shape == (1, 2) and X[:, 1].shape == (3, 1) and X[-1, :].a[0, 0] is X.a[2, 0]
    L = ca.SX.sym("x", 3, 2, 4)
    assert isinstance(L, list) and len(L) == 4 and L[0].shape == (3, 2)
    D = ca.DM.zeros((2, 3))
    D[1, 2] = ca.DM(5.0)
    D[0:2, 0:2] = ca.DM([[1, 2], [3, 4]])
    assert np.array_equal(np.array(D), [[1, 2, 0], [3, 4, 5]])
    assert ca.vertcat(*[[], ca.DM([1, 2]), 3.0]).shape == (3, 1)  # empty lists vanish, scalars become rows


def test_sx_simplifications_decide_the_sparsity(ca):
    x, y = ca.SX.sym("x"), ca.SX.sym("y")
    zero = lambda e: e.a[0, 0].op == "c" and e.a[0, 0].v == 0.0
    assert zero(0.0 * x) and zero(x * 0.0) and zero(x - x) and zero(0.0 / x)
    assert (x + 0.0).a[0, 0] is x.a[0, 0] and (1.0 * x).a[0, 0] is x.a[0, 0] and (x / 1.0).a[0, 0] is x.a[0, 0]
    assert (x * x).a[0, 0].op == "sq" and ((x - y) + y).a[0, 0] is x.a[0, 0] and ((x * y) / y).a[0, 0] is x.a[0, 0]
    # an exact-zero table entry removes the dependency (quirk Q10): mtimes(dense D with a 0.0, X)
    D = ca.DM([[1.0, 0.0], [2.0, 3.0]])
    X = ca.SX.sym("X", 2)
    F = ca.Function("f", [X], [ca.mtimes(D, X)])
    (vals, deps), = F.forward_sparse([np.array([0.5, -1.0])])
    assert sorted(deps[0]) == [0] and sorted(deps[1]) == [0, 1] and np.allclose(vals, [0.5, -2.0])


def test_forward_ad_first_and_second_order_match_finite_differences(ca):
    v = ca.SX.sym("v", 4)
    x, y, z, w = (v[i] for i in range(4))
    f = ca.sin(x * y) / (1.0 + z * z) + ca.exp(-w) * x ** 3 + ca.sqrt(1.5 + y) * ca.atan(z) - ca.cos(w) ** 2 + ca.tanh(x - w)
    F = ca.Function("f", [v], [f])
    p = np.array([0.3, -0.4, 0.7, 0.2])
    (val, grad, hess), = F.forward2_sparse([p])
    fun = lambda q: float(F(q))
    assert abs(val[0] - fun(p)) < 1e-15
    g = np.array([grad[0].get(i, 0.0) for i in range(4)])
    fd = np.array([(fun(p + 1e-6 * e) - fun(p - 1e-6 * e)) / 2e-6 for e in np.eye(4)])
    assert np.allclose(g, fd, rtol=0, atol=1e-8)
    H = np.zeros((4, 4))
    for (i, j), h in hess[0].items():
        H[i, j] = H[j, i] = h
    gfun = lambda q: np.array([F.forward_sparse([q])[0][1][0].get(i, 0.0) for i in range(4)])
    Hfd = np.array([(gfun(p + 1e-6 * e) - gfun(p - 1e-6 * e)) / 2e-6 for e in np.eye(4)])
    assert np.allclose(H, Hfd, rtol=0, atol=1e-7)
    # symbolic gradient (what the reference's Collocation differentiates its Lagrange polynomials with)
    t = ca.SX.sym("t")
    poly = (t - 0.25) * (t + 0.5) * (t - 1.0) / 3.0
    d1 = ca.Function("d", [t], [ca.gradient(poly, t)])
    d2 = ca.Function("d", [t], [ca.gradient(ca.gradient(poly, t), t)])
    assert abs(float(d1(0.4)) - (3 * 0.16 - 2 * 0.75 * 0.4 - 0.375) / 3.0) < 1e-14
    assert abs(float(d2(0.4)) - (6 * 0.4 - 1.5) / 3.0) < 1e-14


def test_integrator_stand_in_is_exact_for_polynomials(ca):
    t = ca.SX.sym("t")
    run = ca.integrator("pint", "idas", {"x": ca.SX.sym("x"), "t": t, "ode": 5.0 * t ** 4 - 3.0 * t ** 2 + 1.0}, {"t0": -1.0, "tf": 0.5})
    exact = (0.5 ** 5 - 0.5 ** 3 + 0.5) - ((-1.0) ** 5 - (-1.0) ** 3 + (-1.0))
    assert abs(float(run(x0=0)["xf"]) - exact) < 1e-14
    assert math.isclose(float(run(x0=2.0)["xf"]), 2.0 + exact, rel_tol=0, abs_tol=1e-14)


def test_kron_diag_solve_as_the_reference_uses_them(ca):
    S = ca.diag(ca.vertcat(np.array([2.0, 4.0])))
    inv = ca.solve(S, np.eye(2))
    assert np.array_equal(np.array(inv), [[0.5, 0.0], [0.0, 0.25]])
    K = ca.kron(ca.DM.eye(2), ca.DM([[1.0, 2.0], [3.0, 4.0]]))
    assert np.array_equal(np.array(K), np.kron(np.eye(2), [[1.0, 2.0], [3.0, 4.0]]))
    x = ca.SX.sym("x", 2)
    y = ca.mtimes(inv, x)  # 0.5 x0 | 0.25 x1: the off-diagonal zeros fold away
    (vals, deps), = ca.Function("f", [x], [y]).forward_sparse([np.array([4.0, 8.0])])
    assert np.allclose(vals, [2.0, 2.0]) and sorted(deps[0]) == [0] and sorted(deps[1]) == [1]
