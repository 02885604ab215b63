"""Pin the oracle NLP: Appendix-A golden, IPOPT-banner size triples from the reference's stored notebook
outputs, finite differences, layout conventions."""
import numpy as np
import pytest

from helpers import random_point
from mpopt_b200 import problems as pr
from oracle.nlp import OracleNLP


def test_golden_G0():
    n = OracleNLP(pr.moon_lander(), 2, [2, 1], "LGR")
    rp, ci = n.structure()
    assert (n.n_z, n.n_g, len(ci)) == (14, 13, 56)
    assert rp.tolist() == [0, 6, 12, 18, 23, 29, 35, 41, 46, 49, 52, 54, 55, 56]
    want = [[0, 1, 2, 4, 12, 13], [0, 1, 2, 5, 12, 13], [0, 1, 2, 6, 12, 13], [2, 3, 7, 12, 13],
            [4, 5, 6, 8, 12, 13], [4, 5, 6, 9, 12, 13], [4, 5, 6, 10, 12, 13], [6, 7, 11, 12, 13],
            [8, 9, 10], [8, 9, 10], [10, 11], [3], [7]]
    assert [ci[rp[r]:rp[r + 1]].tolist() for r in range(13)] == want


# (problem, K, p, scheme) -> (free variables, equalities, inequalities) as printed by IPOPT in the reference docs
BANNERS = [
    ("moon_lander", 10, 6, "LGR", (182, 124, 60)),     # docs/source/notebooks/moon_lander.ipynb:171-176
    ("moon_lander", 2, 30, "CGL", (182, 124, 60)),     # moon_lander.ipynb:314-318
    ("hyper_sensitive", 5, 50, "LGR", (501, 252, 0)),  # hypersensitive.ipynb:165-170
    ("van_der_pol", 1, 25, "LGR", (76, 52, 25)),       # vanderpol.ipynb:177-182
    ("two_phase_schwartz", 1, 20, "LGR", (125, 88, 41)),  # twophaseschwartz.ipynb:195-200
    ("delta3_launch_vehicle", 1, 11, "LGR", (474, 374, 276)),  # multi_stage_launch_vehicle_ascent.ipynb:466-471
    ("falcon9_launcher", 5, 6, "LGR", (956, 746, 641)),       # falcon9_to_orbit.ipynb:480-485 (fork: links (0,1), (0,2))
]


@pytest.mark.parametrize("problem,K,p,scheme,want", BANNERS, ids=[b[0] + str(b[2]) for b in BANNERS])
def test_ipopt_banner_sizes(problem, K, p, scheme, want):
    n = OracleNLP(pr.REGISTRY[problem](), K, p, scheme)
    zmin, zmax, gmin, gmax = n.bounds()
    assert len(zmin) == n.n_z and len(gmin) == n.n_g
    free = int((zmin != zmax).sum())
    eq = int((gmin == gmax).sum())
    assert (free, eq, n.n_g - eq) == want


def test_baseline_config_counts():
    """BASELINE.md section 2 table."""
    n = OracleNLP(pr.moon_lander(), 20, 3, "LGR")
    assert (n.N, n.n_z, n.n_g, len(n.structure()[1])) == (61, 185, 184, 1096)
    po = [30 if k % 3 == 1 else 3 for k in range(2048)]
    assert sum(po) + 1 == 24586


@pytest.mark.parametrize("problem,K,po,scheme", [
    ("kitchen_sink", 3, [3, 2, 4], "LGL"), ("two_phase_schwartz", 2, 4, "LGR"), ("van_der_pol", 3, [2, 5, 3], "CGL"),
    ("robot_arm", 2, 3, "LGR"), ("synthetic_6_3", 2, 4, "LGR")])
def test_jacobian_and_gradient_by_finite_differences(problem, K, po, scheme):
    n = OracleNLP(pr.REGISTRY[problem](), K, po, scheme)
    z, p = random_point(n, dirichlet=True)
    if problem == "robot_arm":
        z = np.abs(z) + 0.5
    J = n.jac_g(z, p).toarray()
    gr = n.grad_f(z, p)
    eps = 1e-6
    for j in range(n.n_z):
        e = np.zeros(n.n_z)
        e[j] = eps
        dg = (n.g(z + e, p) - n.g(z - e, p)) / (2 * eps)
        assert np.abs(dg - J[:, j]).max() < 2e-6 * max(1.0, np.abs(J[:, j]).max()), f"column {j}"
        df = (n.f(z + e, p) - n.f(z - e, p)) / (2 * eps)
        assert abs(df - gr[j]) < 2e-6 * max(1.0, abs(gr[j])), f"grad {j}"


def test_pattern_is_independent_of_the_point():
    n = OracleNLP(pr.kitchen_sink(), 3, [3, 2, 4], "LGL")
    z1, p = random_point(n, seed=1)
    z2, _ = random_point(n, seed=2)
    a, b = n.jac_g(z1, p), n.jac_g(np.zeros_like(z2), p)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


def test_state_major_layout_and_ownership():
    """mpopt.py:537-543 (column-major flatten) and :189-195 (shared node belongs to the earlier segment)."""
    n = OracleNLP(pr.moon_lander(), 3, [2, 3, 1], "LGR")
    assert n.colX(0, 5, 1) == n.N + 5 and n.colU(0, 2, 0) == 2 * n.N + 2
    assert n.node_seg.tolist() == [0, 0, 0, 1, 1, 1, 2] and n.node_loc.tolist() == [0, 1, 2, 1, 2, 3, 1]


def test_initial_guess_and_bounds_layout():
    n = OracleNLP(pr.moon_lander(), 4, 3, "LGR")
    z0 = n.initialize_solution()
    N = n.N
    assert z0.shape == (n.n_z,)
    assert np.allclose(z0[:N], 10.0 + (0.0 - 10.0) * np.linspace(0, 1, N))  # x0 from x00=10 to xf0=0
    assert z0[-2] == 0.0 and z0[-1] == 4.0
    zmin, zmax, gmin, gmax = n.bounds()
    assert zmin[0] == zmax[0] == 10.0 and zmin[N] == zmax[N] == -2.0  # initial state pinned (mpopt.py:550-551)
    assert (zmin[2 * N:3 * N] == 0).all() and (zmax[2 * N:3 * N] == 3).all()
    assert zmin[-1] == 3 and zmax[-1] == 5
    assert (gmin[:2 * N] == 0).all() and (gmax[:2 * N] == 0).all()
    assert (gmin[2 * N:3 * N - 1] == 0).all() and (gmax[2 * N:3 * N - 1] == 3).all()


def test_exact_zero_folding_switch():
    """Q10: with LGL and even p the centre D entry is exactly 0; it leaves the pattern unless d f_s/d x_s is there."""
    ocp = pr.moon_lander()  # d f_0/d x_0 = 0 and d f_1/d x_1 = 0
    a = OracleNLP(ocp, 1, 4, "LGL", drop_exact_zeros=True)
    b = OracleNLP(ocp, 1, 4, "LGL", drop_exact_zeros=False)
    if a.tab.D[4][2, 2] == 0.0:
        assert len(b.structure()[1]) - len(a.structure()[1]) == 2
    else:
        assert len(b.structure()[1]) == len(a.structure()[1])


def test_delta3_banner_bound_types_and_derivatives():
    """The 4-phase launch-vehicle NLP: IPOPT's banner also counts the bound types (all 474 free variables boxed, 132
    two-sided and 144 upper-only inequalities, multi_stage_launch_vehicle_ascent.ipynb:467-474); the Jacobian and the
    gradient of the scaled problem agree with finite differences near the example's initial guess."""
    n = OracleNLP(pr.delta3_launch_vehicle(), 1, 11, "LGR")
    zmin, zmax, gmin, gmax = n.bounds()
    free = zmin != zmax
    assert int((free & np.isfinite(zmin) & np.isfinite(zmax)).sum()) == 474
    ineq = gmin != gmax
    assert int((ineq & np.isfinite(gmin) & np.isfinite(gmax)).sum()) == 132
    assert int((ineq & ~np.isfinite(gmin) & np.isfinite(gmax)).sum()) == 144
    z = n.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(5).standard_normal(n.n_z))
    J = n.jac_g(z).toarray()
    gr = n.grad_f(z)
    for j in np.random.default_rng(6).choice(n.n_z, 60, replace=False):
        e = np.zeros(n.n_z)
        e[j] = 1e-6 * max(1.0, abs(z[j]))
        fd = (n.g(z + e) - n.g(z - e)) / (2 * e[j])
        assert np.abs(fd - J[:, j]).max() < 2e-6 * max(1.0, np.abs(J[:, j]).max()), j
        fdf = (n.f(z + e) - n.f(z - e)) / (2 * e[j])
        assert abs(fdf - gr[j]) < 2e-6 * max(1.0, abs(gr[j])), j


@pytest.mark.parametrize("name,K,po", [("alp_rider", 3, [3, 4, 2]), ("mine_opt", 2, [4, 3]), ("dae_van_der_pol", 3, 3)])
def test_reference_examples_jacobian_by_finite_differences(name, K, po):
    """The remaining problems of tests/test_examples.py:38-50 (time-dependent path row, division by a state, a free
    parameter in a path row)."""
    n = OracleNLP(pr.EXAMPLES[name](), K, po, "LGR")
    z, p = random_point(n, dirichlet=True)
    z = np.abs(z) + 0.5 if name == "mine_opt" else z
    J = n.jac_g(z, p).toarray()
    gr = n.grad_f(z, p)
    eps = 1e-6
    for j in range(n.n_z):
        e = np.zeros(n.n_z)
        e[j] = eps
        fd = (n.g(z + e, p) - n.g(z - e, p)) / (2 * eps)
        assert np.abs(fd - J[:, j]).max() < 1e-6 * max(1.0, np.abs(J[:, j]).max()), j
        assert abs((n.f(z + e, p) - n.f(z - e, p)) / (2 * eps) - gr[j]) < 1e-6 * max(1.0, abs(gr[j])), j
