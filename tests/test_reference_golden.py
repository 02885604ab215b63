"""Parity against THE REFERENCE ITSELF: tests/golden/ref_*.npz hold f, g, grad_f, the CSR Jacobian, the Lagrangian
Hessian, the bounds and the initial guess produced by importing /root/reference/mpopt/mpopt.py unmodified and running
its own ``create_nlp()`` (tests/golden/make_reference_golden.py, oracle/refrun).

* CPU: the oracle reproduces every fixture -- pattern bit for bit (exact-zero folding included), values to 1e-12;
  where the reference tree is present the fixtures are re-derived from it and must come out identical.
* GPU: the CUDA path, through the C ABI, reproduces every fixture -- index arrays bit for bit, values to 1e-10
  (north_star's tolerance).
"""
import os
import sys

import numpy as np
import pytest

from helpers import assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
if GOLD not in sys.path:
    sys.path.insert(0, GOLD)

import make_reference_golden as MR  # noqa: E402

CASES = MR.cases()
NAMES = sorted(CASES)


def _load(name):
    return np.load(os.path.join(GOLD, "ref_" + name + ".npz"))


def test_every_case_has_a_fixture():
    missing = [n for n in NAMES if not os.path.exists(os.path.join(GOLD, "ref_" + n + ".npz"))]
    assert not missing, f"run tests/golden/make_reference_golden.py: {missing}"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_run(name):
    from oracle.adaptive import OracleAdaptiveNLP
    from oracle.hessian import hess_l
    from oracle.nlp import OracleNLP

    fac, K, po, scheme, adaptive = CASES[name][:5]
    G = _load(name)
    ora = (OracleAdaptiveNLP if adaptive else OracleNLP)(fac(), K, po, scheme)  # default mode: exact zeros folded (Q10)
    assert (ora.n_z, ora.n_g) == (G["z"].size, G["g"].size)
    f, g, grad, J = ora._eval(G["z"], G["p"])
    assert np.array_equal(J.indptr, G["rowptr"]) and np.array_equal(J.indices, G["colind"]), "Jacobian pattern"
    assert_close(J.data, G["values"], "jac_g values", 1e-12)
    assert_close(g, G["g"], "g", 1e-12)
    assert_close(grad, G["grad"], "grad_f", 1e-12)
    assert_close(f, float(G["f"]), "f", 1e-12)
    for a, b, what in zip(ora.bounds(), (G["zmin"], G["zmax"], G["gmin"], G["gmax"]), ("lbx", "ubx", "lbg", "ubg")):
        assert np.array_equal(a, b), what
    assert np.array_equal(ora.initialize_solution(), G["z0"]), "initial guess"
    try:
        H = hess_l(ora, G["z"], G["p"], float(G["lam_f"]), G["lam_g"])
    except NotImplementedError:  # widths-as-variables NLP with explicit time dependence: no oracle Hessian
        return
    assert np.array_equal(H.indptr, G["hrowptr"]) and np.array_equal(H.indices, G["hcolind"]), "Hessian pattern"
    assert_close(H.data, G["hvalues"], "hess_l values", 1e-11)


@pytest.mark.parametrize("name", ["moon_lander_K20_p3_LGR", "kitchen_sink_K4_LGR", "moon_all_K3_mixed_CGL",
                                  "schwartz_K5_p10_LGR", "adaptive_sink_K3_LGR", "delta3_K1_p11_LGR"])
def test_fixtures_come_from_the_reference(name):
    """Re-derive a fixture from /root/reference (skipped where the tree does not exist, e.g. on the GPU box)."""
    from oracle.refrun import run_reference as rr

    if not rr.available():
        pytest.skip("reference tree not present")
    out = MR.run(name, CASES[name])
    G = _load(name)
    for k in ("rowptr", "colind", "hrowptr", "hcolind", "zmin", "zmax", "gmin", "gmax", "z0"):
        assert np.array_equal(out[k], G[k]), k
    for k in ("f", "g", "grad", "values", "hvalues"):
        assert_close(out[k], G[k], k, 1e-15)


def test_oracle_goldens_equal_reference_goldens():
    """The fixtures generated from the oracle (make_golden.py) and from the reference at the same points agree."""
    import make_golden as M

    for name in M.CASES:
        A, B = np.load(os.path.join(GOLD, name + ".npz")), _load(name)
        assert np.array_equal(A["z"], B["z"]) and np.array_equal(A["p"], B["p"])
        assert np.array_equal(A["rowptr"], B["rowptr"]) and np.array_equal(A["colind"], B["colind"]), name
        for k in ("g", "values", "grad", "f"):
            assert_close(A[k], B[k], f"{name}: {k}", 1e-12)
        hp = os.path.join(GOLD, name + "_hess.npz")
        if os.path.exists(hp):
            Hh = np.load(hp)
            assert np.array_equal(Hh["rowptr"], B["hrowptr"]) and np.array_equal(Hh["colind"], B["hcolind"]), name
            assert_close(Hh["values"], B["hvalues"], f"{name}: hess_l", 1e-11)


def test_reference_collocation_tables_match_oracle():
    """The reference's own ``Collocation`` (roots, symbolic D, composite W incl. the dropped w[0], interpolation
    matrix) against the oracle's tables, all three schemes."""
    from oracle.refrun import run_reference as rr

    if not rr.available():
        pytest.skip("reference tree not present")
    ref = rr.load_reference()
    from oracle.collocation import Tables

    for scheme in ("LGR", "LGL", "CGL"):
        for deg in (1, 2, 5, 12):
            col = ref.Collocation([deg, deg], scheme)
            tab = Tables([deg], scheme)
            assert_close(col.roots[deg], tab.roots[deg], "roots", 1e-14)
            assert_close(np.array(col.get_diff_matrix(deg)), tab.D[deg], "D", 1e-11)
            assert_close(np.array(col.get_quadrature_weights(deg)).ravel(), tab.w[deg], "w", 1e-13)
            mid = 0.5 * (tab.roots[deg][1:] + tab.roots[deg][:-1])
            assert_close(np.array(col.get_interpolation_matrix(mid, deg)), tab.Cmid[deg], "Cmid", 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_run(libmpx, name):
    """The device keeps every structural entry here (``drop_exact_zeros=False``): which analytically-zero table entries
    come out as exactly 0.0 depends on rounding noise, so the reference's pattern must be CONTAINED in the device's, the
    extra entries must be numerically zero, and everything on the reference's pattern must agree to 1e-10."""
    from mpopt_b200 import _lib
    from mpopt_b200.nlp import Transcription

    fac, K, po, scheme, adaptive = CASES[name][:5]
    G = _load(name)
    tr = Transcription(fac(), K, po, scheme, drop_exact_zeros=False, adaptive=adaptive)
    assert (tr.n_z, tr.n_g) == (G["z"].size, G["g"].size)
    rp, ci = tr.structure()
    g = np.empty(tr.n_g)
    vals = tr.jac_g_values(G["z"], G["p"], g_out=g)
    if np.array_equal(rp, G["rowptr"]) and np.array_equal(ci, G["colind"]):
        assert_close(vals, G["values"], "jac_g values")
    else:
        key = lambda rowptr, col: np.repeat(np.arange(tr.n_g, dtype=np.int64), np.diff(rowptr)) * tr.n_z + col
        kd, kr = key(rp, ci), key(G["rowptr"], G["colind"])  # both sorted: CSR with sorted columns
        pos = np.searchsorted(kd, kr)
        assert (pos < kd.size).all() and np.array_equal(kd[np.minimum(pos, kd.size - 1)], kr), \
            "reference entries missing from the device pattern"
        only_dev = np.ones(kd.size, bool)
        only_dev[pos] = False
        assert np.abs(vals[only_dev]).max(initial=0.0) <= 1e-12, "device-only entries are not numerical zeros"
        assert_close(vals[pos], G["values"], "jac_g values")
    assert_close(g, G["g"], "g")
    assert_close(tr.grad_f(G["z"], G["p"]), G["grad"], "grad_f")
    assert_close(tr.f(G["z"], G["p"]), float(G["f"]), "f")
    for a, b, what in zip(tr.bounds(), (G["zmin"], G["zmax"], G["gmin"], G["gmax"]), ("lbx", "ubx", "lbg", "ubg")):
        assert np.array_equal(a, b), what
    assert np.allclose(tr.initial_guess(), G["z0"], rtol=0, atol=1e-15)
    try:
        hrp, hci = tr.hess_structure()
    except _lib.MpxError:
        assert adaptive  # only the widths-as-variables NLP with explicit time dependence is refused
        return
    assert np.array_equal(hrp, G["hrowptr"]) and np.array_equal(hci, G["hcolind"]), "Hessian pattern"
    assert_close(tr.hess_l_values(G["z"], G["p"], float(G["lam_f"]), G["lam_g"]), G["hvalues"], "hess_l values")


# ---------------------------------------------------------------------------------------------------------------------
# interpolation / residual path: fixtures from the reference's own residual functions

import make_reference_residual_golden as MRR  # noqa: E402


def _split(flat, counts):
    off = np.concatenate([[0], np.cumsum(counts)])
    return [flat[off[k]: off[k + 1]] for k in range(len(counts))]


@pytest.mark.parametrize("name", sorted(MRR.CASES))
def test_oracle_residual_path_matches_reference_run(name):
    from mpopt_b200.problems import REGISTRY
    from oracle import residual as R
    from oracle.nlp import OracleNLP

    problem, K, deg, scheme, _ = MRR.CASES[name]
    G = np.load(os.path.join(GOLD, "refres_" + name + ".npz"))
    ora = OracleNLP(REGISTRY[problem](), K, deg, scheme)
    z, p = G["z"], G["p"]
    for ph in range(int(G["n_phases"])):
        for grid in MRR.GRIDS:
            key = f"ph{ph}_{grid}_"
            taus = R.residual_grid_taus(ora, ph, grid, p)
            assert [len(t) for t in taus] == G[key + "counts"].tolist(), "points per segment"
            assert_close(np.concatenate(taus), G[key + "taus"], "target points", 1e-13)
            taus = _split(G[key + "taus"], G[key + "counts"])
            Xi, Ui, ti, DXi, DUi = R.interpolate_phase(ora, z, p, ph, taus)
            for a, k in ((Xi, "xi"), (Ui, "ui"), (ti, "ti"), (DXi, "dxi"), (DUi, "dui")):
                assert_close(a, G[key + k], k, 1e-12)
            assert_close(R.dynamics_residuals_phase(ora, z, p, ph, taus)[1], G[key + "res"], "residual", 1e-12)
            _, ddx, ddu = R.second_derivatives_phase(ora, z, p, ph, taus)
            assert_close(ddx, G[key + "ddxi"], "ddxi", 1e-11)
            assert_close(ddu, G[key + "ddui"], "ddui", 1e-11)
            xint, resx, _ = R.states_from_dynamics_phase(ora, z, p, ph, taus)
            assert_close(xint, G[key + "xint"], "xint", 1e-12)
            assert_close(resx, G[key + "res_x"], "res_x", 1e-12)


def test_residual_fixtures_come_from_the_reference():
    from oracle.refrun import run_reference as rr

    if not rr.available():
        pytest.skip("reference tree not present")
    name = "sink_K3_p5_LGL"
    out, G = MRR.run(name), np.load(os.path.join(GOLD, "refres_" + name + ".npz"))
    assert sorted(out) == sorted(G.files)
    for k in G.files:
        assert_close(np.asarray(out[k], float), np.asarray(G[k], float), k, 1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MRR.CASES))
def test_cuda_residual_path_matches_reference_run(libmpx, name):
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY

    problem, K, deg, scheme, _ = MRR.CASES[name]
    G = np.load(os.path.join(GOLD, "refres_" + name + ".npz"))
    tr = Transcription(REGISTRY[problem](), K, deg, scheme)
    z, p = G["z"], G["p"]
    for ph in range(int(G["n_phases"])):
        for grid in MRR.GRIDS:
            key = f"ph{ph}_{grid}_"
            taus = _split(G[key + "taus"], G[key + "counts"])
            r = tr.residuals(z, p, phase=ph, taus=taus)
            for k in ("xi", "ui", "ti", "dxi", "dui", "res"):
                assert_close(r[k], G[key + k], f"{grid} {k}")
            d2 = tr.second_derivatives(z, p, phase=ph, taus=taus)
            assert_close(d2["ddxi"], G[key + "ddxi"], f"{grid} ddxi", 1e-9)
            assert_close(d2["ddui"], G[key + "ddui"], f"{grid} ddui", 1e-9)
            s = tr.state_residuals(z, p, phase=ph, taus=taus)
            assert_close(s["xint"], G[key + "xint"], f"{grid} xint")
            assert_close(s["res_x"], G[key + "res_x"], f"{grid} res_x")


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own solve() against mp.solve of this package

def _solutions():
    import json

    with open(os.path.join(GOLD, "ref_solutions.json")) as f:
        return json.load(f)


@pytest.mark.gpu
@pytest.mark.parametrize("rec", _solutions(), ids=lambda r: f"{r['cls']}-{r['problem']}-{r['n_segments']}x{r['poly_orders']}-{r['scheme']}")
def test_mp_solve_matches_reference_solve(libmpx, rec):
    """tests/golden/ref_solutions.json: optima of the UNMODIFIED reference's ``mpopt(...).solve()`` /
    ``mpopt_adaptive(...).solve()`` (run on the stand-ins of oracle/refrun; BASELINE.json's config 1 is the first record).
    The same call through ``mpopt_b200.mp`` -- GPU evaluators behind the same interior-point method -- must land on the
    same optimum."""
    from mpopt_b200 import mp
    from mpopt_b200.problems import REGISTRY

    assert rec["success"]
    mp.mpopt._MUTE_ = True
    mpo = getattr(mp, rec["cls"])(REGISTRY[rec["problem"]](), rec["n_segments"], rec["poly_orders"], rec["scheme"])
    sol = mpo.solve(nlp_solver_options={"ipopt.tol": 1e-10})
    f = float(sol["f"])
    if abs(rec["f"]) < 1e-12:
        assert abs(f) < 1e-10
    else:
        rtol = 1e-6 if rec["cls"] == "mpopt" else 1e-4  # the widths-as-variables NLP is not convex: same basin, looser
        assert abs(f - rec["f"]) <= rtol * abs(rec["f"]), f"{f!r} vs the reference's {rec['f']!r}"
