"""Pin the oracle of the ``mpopt_adaptive`` NLP (widths as variables): row / variable counts from
the reference's formulas (mpopt.py:2927-2979, :3034-3136, :3169), finite differences of g and f, bounds."""
import numpy as np
import pytest

from mpopt_b200 import problems as pr
from oracle.adaptive import OracleAdaptiveNLP
from oracle.nlp import OracleNLP


def adaptive_point(n, seed=20261017):
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, n.n_z)
    for ph in range(n.P):
        z[n.colT0(ph)] = 0.25 * ph + (0.1 if ph else 0.0)
        z[n.colTF(ph)] = 2.0 + 1.5 * ph
        for m in range(n.na):
            z[n.colA(ph, m)] = rng.uniform(0.2, 1.2)
        z[n.colW(ph, np.arange(n.K))] = rng.dirichlet(np.ones(n.K)) * 0.8 + 0.2 / n.K
    return z


def test_sizes_moon_lander():
    """tests/test_mpopt.py:258-259 builds mpopt_adaptive(moon_lander, 3, 3): N = 10."""
    n = OracleAdaptiveNLP(pr.moon_lander(), 3, 3, "LGR")
    N, K = 10, 3
    assert n.n_z == N * 3 + 2 + K and n.n_p == 0
    # F 2N | TC 2 | sum 1 | ui (N-1) | xi 2(N-1) | residuals 2(N-1)
    assert n.n_g == 2 * N + 2 + 1 + (N - 1) + 2 * (N - 1) + 2 * (N - 1)
    n2 = OracleAdaptiveNLP(pr.moon_lander(), 3, 3, "LGR", mid_residuals=False)   # tests/test_mpopt.py:474
    assert n2.n_g == n.n_g - 2 * (N - 1)
    zmin, zmax, gmin, gmax = n.bounds()
    assert len(zmin) == len(zmax) == n.n_z and len(gmin) == len(gmax) == n.n_g
    assert (zmin[-K:] == 1e-4).all() and (zmax[-K:] == 1.0).all()
    assert (gmin[-2 * (N - 1):] == -1e-3).all() and (gmax[-2 * (N - 1):] == 1e-3).all()
    z0 = n.initialize_solution()
    assert len(z0) == n.n_z and np.allclose(z0[-K:], 1 / K)


def test_reduces_to_fixed_width_nlp():
    """With the widths frozen, F / C / DU / TC rows and the objective equal those of the base NLP with p = w."""
    ocp = pr.kitchen_sink()
    a = OracleAdaptiveNLP(ocp, 3, [3, 2, 4], "LGR")
    b = OracleNLP(ocp, 3, [3, 2, 4], "LGR")
    z = adaptive_point(a)
    nb = b.nvar
    zb = np.concatenate([z[ph * a.nvar: ph * a.nvar + nb] for ph in range(a.P)])
    p = np.concatenate([z[a.colW(ph, np.arange(a.K))] for ph in range(a.P)])
    ga, gb = a.g(z), b.g(zb, p)
    for ph in range(a.P):
        Ra, Rb = a._rows[ph], b._rows[ph]
        for key, n_rows in (("F", a.nx * a.N), ("C", Ra["nc"] * a.N), ("DU", a.nu * a.N), ("TC", Ra["ntc"])):
            ra, rb = a.row_off[ph] + Ra[key], b.row_off[ph] + Rb[key]
            assert np.array_equal(ga[ra: ra + n_rows], gb[rb: rb + n_rows]), key
    assert a.f(z) == b.f(zb, p)


@pytest.mark.parametrize("problem,K,po,scheme", [
    ("moon_lander", 3, 3, "LGR"), ("hyper_sensitive", 3, [4, 2, 3], "LGL"), ("kitchen_sink", 3, [3, 2, 4], "LGR"),
    ("two_phase_schwartz", 2, 4, "CGL"), ("synthetic_6_3", 2, 3, "LGR")])
def test_jacobian_and_gradient_by_finite_differences(problem, K, po, scheme):
    n = OracleAdaptiveNLP(pr.REGISTRY[problem](), K, po, scheme)
    z = adaptive_point(n)
    J = n.jac_g(z).toarray()
    gr = n.grad_f(z)
    eps = 1e-6
    for j in range(n.n_z):
        e = np.zeros(n.n_z)
        e[j] = eps
        fd = (n.g(z + e) - n.g(z - e)) / (2 * eps)
        assert np.abs(fd - J[:, j]).max() < 5e-7 * max(1.0, np.abs(J[:, j]).max()), (j, np.abs(fd - J[:, j]).argmax())
        fdf = (n.f(z + e) - n.f(z - e)) / (2 * eps)
        assert abs(fdf - gr[j]) < 5e-7 * max(1.0, abs(gr[j])), j


def test_structure_is_point_independent_and_has_no_spurious_entries():
    n = OracleAdaptiveNLP(pr.kitchen_sink(), 3, [3, 2, 4], "LGR")
    z1, z2 = adaptive_point(n, 1), adaptive_point(n, 2)
    z1[n.colT0(0)], z2[n.colT0(0)] = 0.3, 0.4   # t0 = 0 would zero sin(t) u / a t at node 0 by coincidence
    J1, J2 = n.jac_g(z1), n.jac_g(z2)
    assert np.array_equal(J1.indptr, J2.indptr) and np.array_equal(J1.indices, J2.indices)
    # every structural entry is non-zero at (at least one of) two generic points
    assert ((J1.data != 0) | (J2.data != 0)).all()


def test_host_layout_agrees_with_oracle():
    """The Python twin of the plan's layout (mpopt_b200/layout.py, no GPU): variables, rows and the width columns of
    the adaptive NLP sit where the oracle puts them."""
    from mpopt_b200.layout import Layout
    from mpopt_b200.program import Program

    for name, K, po in (("moon_lander", 3, [3, 3, 3]), ("kitchen_sink", 3, [3, 2, 4]), ("falcon9_launcher", 2, [5, 3])):
        ocp = pr.REGISTRY[name]()
        ora = OracleAdaptiveNLP(ocp, K, po, "LGR")
        sw_u = [ora._rows[ph]["sw_u"] for ph in range(ora.P)]
        sw_x = [ora._rows[ph]["sw_x"] for ph in range(ora.P)]
        L = Layout(Program(ocp), po, [bool(v) for v in ocp.diff_u], [False] * ora.P, [False] * ora.P,
                   ora.n_links, adaptive=dict(sw_u=sw_u, sw_x=sw_x, mid_residuals=True))
        assert (L.n_z, L.n_p, L.n_g) == (ora.n_z, 0, ora.n_g)
        for ph in range(ora.P):
            assert L.phases[ph].gSW == ora.row_off[ph] + ora._rows[ph]["SW"]
            assert L.phases[ph].gTC == ora.row_off[ph] + ora._rows[ph]["TC"]
            assert L.colW(ph, K - 1) == ora.colW(ph, K - 1) and L.colT0(ph) == ora.colT0(ph)


@pytest.mark.parametrize("problem,K,po", [("moon_lander", 3, 3), ("hyper_sensitive", 3, [4, 2, 3]), ("synthetic_6_3", 2, 3),
                                          ("two_phase_schwartz", 2, 4)])
def test_adaptive_hessian_oracle_by_finite_differences(problem, K, po):
    """Hessian of lam_f f + lam_g . g of the widths-as-variables NLP (the nlp_hess_l CasADi would derive for
    mpopt_adaptive, mpopt.py:3174-3205): second-order duals with the widths as variables and the dense per-segment
    coupling of the mid-point residual rows, against central differences of the oracle's own gradient of the Lagrangian.
    (Head start for the device kernel: DESIGN.md section 8, item 1.)"""
    import scipy.sparse as sp
    from oracle.hessian import hess_l

    n = OracleAdaptiveNLP(pr.REGISTRY[problem](), K, po, "LGR")
    z = adaptive_point(n)
    rng = np.random.default_rng(5)
    lam, sig = rng.uniform(-1, 1, n.n_g), 0.7
    H = hess_l(n, z, None, sig, lam)
    assert (H.tocoo().row >= H.tocoo().col).all()
    Hs = (H + sp.tril(H, k=-1).T).toarray()
    pattern = (H + sp.tril(H, k=-1).T).astype(bool).toarray()

    def grad_lag(zz):
        return sig * n.grad_f(zz) + n.jac_g(zz).T @ lam

    e = 1e-6
    for j in range(n.n_z):
        dz = np.zeros(n.n_z)
        dz[j] = e
        col = (grad_lag(z + dz) - grad_lag(z - dz)) / (2 * e)
        assert np.abs(col - Hs[:, j]).max() <= 1e-6 * max(1.0, np.abs(col).max()), j
        assert np.abs(col[~pattern[:, j]]).max(initial=0.0) <= 1e-6, j   # nothing outside the structural pattern


def test_adaptive_hessian_oracle_refuses_time_dependence():
    from oracle.hessian import hess_l

    n = OracleAdaptiveNLP(pr.kitchen_sink(), 3, [3, 2, 4], "LGR")
    with pytest.raises(NotImplementedError):
        hess_l(n, adaptive_point(n), None, 1.0, np.zeros(n.n_g))
