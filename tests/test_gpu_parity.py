"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bit-exact on the CSR structure, 1e-10 relative on float64 values (north_star's tolerance)."""
import numpy as np
import pytest

from helpers import assert_close, random_point

pytestmark = pytest.mark.gpu

CASES = [
    # name, problem, K, poly_orders, scheme, dirichlet widths
    ("moon_cfg1", "moon_lander", 20, 3, "LGR", False),
    ("moon_G0", "moon_lander", 2, [2, 1], "LGR", False),
    ("moon_p15", "moon_lander", 64, 15, "LGR", True),
    ("moon_diffu", "moon_lander", 6, 4, "LGL", True),
    ("hyper", "hyper_sensitive", 15, 15, "LGR", False),
    ("vdp_mixed_cgl", "van_der_pol", 48, [30 if k % 3 == 1 else 3 for k in range(48)], "CGL", True),
    ("vdp_lgl", "van_der_pol", 1, 15, "LGL", False),
    ("schwartz", "two_phase_schwartz", 5, 10, "LGR", True),
    ("schwartz_lgl", "two_phase_schwartz", 1, 15, "LGL", False),
    ("generic2", "generic_two_phase", 3, [2, 5, 3], "CGL", True),
    ("robot", "robot_arm", 20, 4, "LGR", False),
    ("syn63", "synthetic_6_3", 37, 15, "LGR", True),
    ("syn63_lgl20", "synthetic_6_3", 9, 20, "LGL", False),
    ("sink", "kitchen_sink", 7, [3, 4, 6, 2, 5, 4, 1], "LGL", True),
    ("sink_p1", "kitchen_sink", 4, 1, "LGR", False),
    ("p30_single", "moon_lander", 1, 30, "CGL", False),
    ("hyper_p50", "hyper_sensitive", 5, 50, "LGR", False),   # docs/source/notebooks/hypersensitive.ipynb:165-170 (v1 kernel: degree > 31)
    ("vdp_p25", "van_der_pol", 1, 25, "LGR", False),          # vanderpol.ipynb:177-182
    ("moon_p40_mixed", "moon_lander", 3, [40, 6, 33], "LGL", True),
    ("delta3_p11", "delta3_launch_vehicle", 1, 11, "LGR", False),   # multi_stage_launch_vehicle_ascent.ipynb:466-471
    ("delta3_mixed", "delta3_launch_vehicle", 4, [4, 6, 5, 3], "LGL", True),
    ("falcon9_5x6", "falcon9_launcher", 5, 6, "LGR", False),        # falcon9_to_orbit.ipynb:480-485: fork links, constant path rows
    ("falcon9_mixed", "falcon9_launcher", 3, [7, 2, 5], "CGL", True),
]


def _ocp(problem, scheme):
    from mpopt_b200.problems import REGISTRY

    ocp = REGISTRY[problem]()
    if problem == "moon_lander" and scheme == "LGL":
        ocp.diff_u[0], ocp.du_continuity[0] = 1, 1
    return ocp


def _build(problem, K, po, scheme, drop=True, device_tables=False):
    """(oracle, Transcription).  With ``device_tables`` the oracle is handed the plan's own tables (checked
    against the oracle's to <= 1e-11 first) so that the exact-zero folding (Q10), which hinges on rounding
    noise in analytically-zero entries, is decided on identical numbers."""
    from mpopt_b200.nlp import Transcription
    from oracle.nlp import OracleNLP

    ocp = _ocp(problem, scheme)
    tr = Transcription(ocp, K, po, scheme, drop_exact_zeros=drop)
    ora = OracleNLP(ocp, K, po, scheme, drop_exact_zeros=drop)
    if device_tables:
        tabs = {}
        for d in sorted(set(tr.poly_orders)):
            r, D, w, Cm = tr.tables(d)
            assert_close(r, ora.tab.roots[d], f"roots[{d}]", 1e-13)
            assert_close(D, ora.tab.D[d], f"D[{d}]", 1e-11)
            assert_close(w, ora.tab.w[d], f"w[{d}]", 1e-13)
            assert_close(Cm, ora.tab.Cmid[d], f"Cmid[{d}]", 1e-12)
            tabs[d] = (r, D, w, Cm)
        ora = OracleNLP(ocp, K, po, scheme, drop_exact_zeros=drop, tables=tabs)
    return ora, tr


def _point(ora, problem, dirichlet):
    z, p = random_point(ora, dirichlet=dirichlet)
    if problem == "robot_arm":
        z = np.abs(z) + 0.5  # keep sin(x4) and the inertia terms away from zero
    if problem in ("delta3_launch_vehicle", "falcon9_launcher"):  # stay near the ascent trajectory: |r| ~ Re, acos arguments inside (-1, 1)
        z = ora.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(7).standard_normal(ora.n_z))
    return z, p


@pytest.mark.parametrize("name,problem,K,po,scheme,dirichlet", CASES, ids=[c[0] for c in CASES])
def test_g_jac_f_grad_match_oracle(libmpx, name, problem, K, po, scheme, dirichlet):
    """Fully independent comparison: oracle tables from scipy, every structural entry kept."""
    ora, tr = _build(problem, K, po, scheme, drop=False)
    assert (tr.n_z, tr.n_p, tr.n_g) == (ora.n_z, ora.n_p, ora.n_g)
    assert tr.program_origin.startswith("aot:")
    z, p = _point(ora, problem, dirichlet)
    rp, ci = tr.structure()
    J = ora.jac_g(z, p)
    assert np.array_equal(rp, J.indptr.astype(np.int64)), "rowptr differs"
    assert np.array_equal(ci, J.indices.astype(np.int64)), "colind differs"
    g = np.empty(tr.n_g)
    vals = tr.jac_g_values(z, p, g_out=g)
    assert_close(g, ora.g(z, p), "g (fused)")
    assert_close(vals, J.data, "jac_g values")
    assert_close(tr.g(z, p), ora.g(z, p), "g")
    assert_close(tr.f(z, p), ora.f(z, p), "f")
    assert_close(tr.grad_f(z, p), ora.grad_f(z, p), "grad_f")
    # a second, nearby point: nothing is cached between evaluations
    z2 = z + 1e-3 * np.random.default_rng(1).standard_normal(z.size)
    assert_close(tr.jac_g_values(z2, p), ora.jac_g(z2, p).data, "jac_g values (2nd point)")


@pytest.mark.parametrize("name,problem,K,po,scheme,dirichlet", CASES, ids=[c[0] for c in CASES])
def test_folded_pattern_matches_oracle(libmpx, name, problem, K, po, scheme, dirichlet):
    """Default mode (exact-zero table entries folded away like CasADi's SX does): tables agree with the
    oracle's, then structure is bit-exact and values agree on the folded pattern."""
    ora, tr = _build(problem, K, po, scheme, drop=True, device_tables=True)
    z, p = _point(ora, problem, dirichlet)
    rp, ci = tr.structure()
    J = ora.jac_g(z, p)
    assert np.array_equal(rp, J.indptr.astype(np.int64)), "rowptr differs"
    assert np.array_equal(ci, J.indices.astype(np.int64)), "colind differs"
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "jac_g values")
    assert_close(g, ora.g(z, p), "g")


def test_golden_G0_structure(libmpx):
    """SURVEY.md Appendix A golden: moon-lander, LGR, K=2, poly_orders=[2,1]."""
    _, tr = _build("moon_lander", 2, [2, 1], "LGR", drop=False)
    rp, ci = tr.structure()
    assert (tr.n_z, tr.n_g, tr.nnz) == (14, 13, 56)
    assert rp.tolist() == [0, 6, 12, 18, 23, 29, 35, 41, 46, 49, 52, 54, 55, 56]
    rows = [ci[rp[r]:rp[r + 1]].tolist() for r in range(13)]
    assert rows[0] == [0, 1, 2, 4, 12, 13] and rows[3] == [2, 3, 7, 12, 13] and rows[7] == [6, 7, 11, 12, 13]
    assert rows[8] == [8, 9, 10] and rows[10] == [10, 11] and rows[11] == [3] and rows[12] == [7]


def test_ccs_adapter(libmpx):
    ora, tr = _build("two_phase_schwartz", 4, 5, "LGR", device_tables=True)
    z, p = random_point(ora)
    cp, ri, perm = tr.structure_ccs()
    Jc = ora.jac_g(z, p).tocsc()
    Jc.sort_indices()
    assert np.array_equal(cp, Jc.indptr) and np.array_equal(ri, Jc.indices)
    assert_close(tr.jac_g_values(z, p)[perm], Jc.data, "CCS values")


def test_exact_zero_folding_changes_pattern(libmpx):
    """LGL interior diagonal entries are analytically 0; whichever come out exactly 0.0 leave the pattern."""
    ora, tr = _build("van_der_pol", 3, 4, "LGL", drop=True, device_tables=True)
    _, tr_full = _build("van_der_pol", 3, 4, "LGL", drop=False)
    D = tr.tables(4)[1]
    assert tr.nnz <= tr_full.nnz
    if D[2, 2] == 0.0:  # centre node of a symmetric 5-point rule
        assert tr.nnz < tr_full.nnz


def test_config2_size_properties(libmpx):
    """BASELINE config 2 at full size through size-independent properties: counts, linearity of the constant
    blocks, row sums of D (derivative of a constant is 0) and agreement with the oracle on a sampled sub-range."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import moon_lander

    tr = Transcription(moon_lander(), 4096, 15, "LGR")
    assert (tr.N, tr.n_z, tr.n_g, tr.nnz) == (61441, 184325, 184324, 3317800)
    rng = np.random.default_rng(5)
    z = rng.uniform(-1, 1, tr.n_z)
    z[-2:] = [0.0, 4.0]
    N = tr.N
    zc = z.copy()
    zc[:2 * N] = 3.25  # constant states: D.X = 0 so F = -h f
    g = tr.g(zc)
    h = (4.0 / 2.0) * (1.0 / 4096)
    assert_close(g[:N], -h * 3.25 * np.ones(N), "F rows of state 0 on constant states", 1e-9)
    # mid-point rows are linear in U
    g1, g2 = tr.g(z), tr.g(2.0 * z)
    mu = slice(2 * N, 2 * N + N - 1)
    assert_close(g2[mu], 2.0 * g1[mu], "mid-point rows linear in U", 1e-12)


def test_headline_counts(libmpx):
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import synthetic_6_3

    tr = Transcription(synthetic_6_3(), 4096, 15, "LGR")
    assert (tr.n_z, tr.n_g, tr.nnz) == (552971, 552966, 12533916)
    from oracle.nlp import OracleNLP

    tabs = {15: tr.tables(15)}
    ora = OracleNLP(synthetic_6_3(), 4096, 15, "LGR", tables=tabs)
    z, p = random_point(ora, tf=1.0, dirichlet=True)
    J = ora.jac_g(z, p)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "headline jac values")
    assert_close(g, ora.g(z, p), "headline g")


def _full_size_against_oracle(make, K, po, scheme, counts, tf):
    """Whole problem against the oracle (structure bit-exact, values 1e-10); the oracle gets the device's tables, which
    the table tests hold to <= 1e-11 of its own, so that exact-zero folding picks the same entries."""
    from mpopt_b200.nlp import Transcription
    from oracle.nlp import OracleNLP

    tr = Transcription(make(), K, po, scheme)
    assert (tr.n_z, tr.n_g, tr.nnz) == counts
    degs = sorted(set(po)) if isinstance(po, list) else [po]
    ora = OracleNLP(make(), K, po, scheme, tables={d: tr.tables(d) for d in degs})
    z, p = random_point(ora, tf=tf, dirichlet=True)
    J = ora.jac_g(z, p)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "jac_g values")
    assert_close(g, ora.g(z, p), "g")
    assert_close(tr.f(z, p), ora.f(z, p), "f")
    assert_close(tr.grad_f(z, p), ora.grad_f(z, p), "grad_f")
    return tr


def test_config3_full_size(libmpx):
    """BASELINE config 3: van-der-Pol, 2048 segments of mixed degree [3, 30, 3, ...], CGL."""
    from mpopt_b200.problems import van_der_pol

    po = [30 if k % 3 == 1 else 3 for k in range(2048)]
    tr = _full_size_against_oracle(van_der_pol, 2048, po, "CGL", (73760, 73757, 2126820), 10.0)
    assert "gjac=v2" in tr.program_origin and "/d" not in tr.program_origin  # mixed degrees: generic instance


def test_config4_full_size(libmpx):
    """BASELINE config 4: synthetic 6/3, 8192 segments of degree 20, LGL (odd block offsets: thread-store fallback)."""
    from mpopt_b200.problems import synthetic_6_3

    tr = _full_size_against_oracle(synthetic_6_3, 8192, 20, "LGL", (1474571, 1474566, 40796346), 1.0)
    assert tr.program_origin.endswith("gjac=v2/d20")


def test_config5_full_size(libmpx):
    """BASELINE config 5 stand-in: two-phase Schwartz, 1024 segments per phase, degree 10, LGR."""
    from mpopt_b200.problems import two_phase_schwartz

    from mpopt_b200.nlp import Transcription
    from oracle.nlp import OracleNLP

    tr = Transcription(two_phase_schwartz(), 1024, 10, "LGR")
    ora = OracleNLP(two_phase_schwartz(), 1024, 10, "LGR", tables={10: tr.tables(10)})
    assert (tr.n_z, tr.n_g, tr.nnz) == (ora.n_z, ora.n_g, ora.jac_g(*random_point(ora)).nnz)
    z, p = random_point(ora, dirichlet=True)
    J = ora.jac_g(z, p)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "jac_g values")
    assert_close(g, ora.g(z, p), "g")
    assert tr.program_origin.endswith("gjac=v2/d10;phases=fused")
    l0 = tr.launches
    tr.jac_g_values(z, p, g_out=g)
    assert tr.launches - l0 == 1, "both phases and the phase-link rows are ONE launch"


@pytest.mark.parametrize("env", [{"MPX_JIT": "1"}, {"MPX_NOSPEC": "1"}, {"MPX_KERNEL": "v4"}, {"MPX_KERNEL": "v4", "MPX_NOSPEC": "1"},
                                 {"MPX_KERNEL": "v1"}, {"MPX_CONST_FIRST": "0", "MPX_V2_NBUF": "1", "MPX_PDL": "0"}])
def test_kernel_variants_agree(libmpx, monkeypatch, env):
    """Every g + jac_g kernel variant (degree-specialised AOT / NVRTC / generic v2, row-block teams v4, CTA-per-segment
    v1, and v2 with its scheduling features off) agrees with the default on the same inputs to 1e-13."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import kitchen_sink, synthetic_6_3

    for make, K, po in ((synthetic_6_3, 40, 15), (kitchen_sink, 9, 6), (synthetic_6_3, 11, 4)):
        ref = Transcription(make(), K, po, "LGR")
        rng = np.random.default_rng(11)
        z = rng.uniform(-1, 1, ref.n_z)
        nvar = ref.n_z // ref.P
        for ph in range(ref.P):
            z[(ph + 1) * nvar - 2 - ref.na] = 0.3 * ph
            z[(ph + 1) * nvar - 1 - ref.na] = 1.5 + ph
        w = np.concatenate([rng.dirichlet(np.ones(K)) for _ in range(ref.P)])
        g0 = np.empty(ref.n_g)
        v0 = ref.jac_g_values(z, w, g_out=g0)
        go0 = ref.g(z, w)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        tr = Transcription(make(), K, po, "LGR")
        for k in env:
            monkeypatch.delenv(k)
        if env.get("MPX_JIT") == "1" and "/d" not in ref.program_origin:
            assert "/jit-d" in tr.program_origin
        if env.get("MPX_KERNEL") in ("v4", "v1"):
            assert "gjac=" + env["MPX_KERNEL"] in tr.program_origin
        g1 = np.empty(tr.n_g)
        v1 = tr.jac_g_values(z, w, g_out=g1)
        # different instances may contract a*b+c differently: a few ulp, far inside the 1e-10 parity tolerance
        assert_close(v1, v0, f"values {env} {tr.program_origin}", 1e-13)
        assert_close(g1, g0, f"g {env} {tr.program_origin}", 1e-13)
        assert_close(tr.g(z, w), go0, f"g-only {env}", 1e-13)


def test_unregistered_problem_compiles_at_run_time(libmpx):
    """A problem that is not in problems.REGISTRY: its functors are compiled through NVRTC from the same kernel
    header, and the result has the same parity with the oracle."""
    from mpopt_b200 import OCP, ca
    from mpopt_b200.nlp import Transcription
    from oracle.nlp import OracleNLP

    ocp = OCP(n_states=3, n_controls=2, n_params=1)
    ocp.dynamics[0] = lambda x, u, t, a: [x[1] * ca.cos(x[2]) + a[0], u[0] * x[0] - ca.exp(-t) * x[1], u[1] / (2.0 + x[0] ** 2)]
    ocp.path_constraints[0] = lambda x, u, t, a: [x[0] * x[0] + u[1] - 4.0]
    ocp.running_costs[0] = lambda x, u, t, a: u[0] * u[0] + ca.sqrt(1.0 + u[1] * u[1]) + a[0] * x[2]
    ocp.terminal_constraints[0] = lambda xf, tf, x0, t0, a: [xf[0] - 1.0, xf[1] * a[0]]
    ocp.terminal_costs[0] = lambda xf, tf, x0, t0, a: tf + xf[2] ** 2
    ocp.lbu[0], ocp.ubu[0] = [-1.0, -2.0], [1.0, 2.0]
    ocp.a0[0] = [0.5]
    ocp.validate()
    K, po = 6, [4, 7, 3, 7, 4, 5]
    tr = Transcription(ocp, K, po, "LGR", drop_exact_zeros=False)
    assert tr.program_origin.startswith("nvrtc:")
    ora = OracleNLP(ocp, K, po, "LGR", drop_exact_zeros=False)
    z, p = random_point(ora, dirichlet=True)
    J = ora.jac_g(z, p)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "jac_g values (nvrtc)")
    assert_close(g, ora.g(z, p), "g (nvrtc)")
    assert_close(tr.f(z, p), ora.f(z, p), "f (nvrtc)")
    assert_close(tr.grad_f(z, p), ora.grad_f(z, p), "grad_f (nvrtc)")
    tr2 = Transcription(ocp, 3, 5, "CGL")  # second plan of the same program: served from the in-process cache
    assert tr2.program_origin == tr.program_origin.split(";")[0] + ";" + tr2.program_origin.split(";")[1]


@pytest.mark.gpu
def test_peer_stores_replicate_every_output():
    """mpx_eval_g_jac_dev_peers (fused evaluation + all-gather): every g / Jacobian value also lands in the peers'
    buffers.  One GPU is enough to check the replication: the 'peers' are two more buffer pairs on the same device;
    a shard plan must leave the rows of other shards untouched."""
    import torch

    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import synthetic_6_3, two_phase_schwartz

    for make, K, p, seg in ((synthetic_6_3, 12, 15, (4, 9)), (synthetic_6_3, 7, 4, None), (two_phase_schwartz, 6, 5, (2, 6))):
        ocp = make()
        tr = Transcription(ocp, K, p, "LGR", drop_exact_zeros=False, segments=seg)
        rng = np.random.default_rng(3)
        z = rng.uniform(-1, 1, tr.n_z)
        for ph in range(tr.P):
            z[(ph + 1) * (tr.n_z // tr.P) - 2 - tr.na] = 0.5 * ph
            z[(ph + 1) * (tr.n_z // tr.P) - 1 - tr.na] = 2.0 + ph
        w = np.concatenate([rng.dirichlet(np.ones(K)) for _ in range(tr.P)])
        dev = torch.device("cuda", 0)
        zd, wd = torch.from_numpy(z).to(dev), torch.from_numpy(w).to(dev)
        mk = lambda n: torch.full((n,), -7.0, dtype=torch.float64, device=dev)
        g0, v0 = mk(tr.n_g), mk(tr.nnz)
        tr.g_jac_dev(zd.data_ptr(), wd.data_ptr(), g0.data_ptr(), v0.data_ptr())
        tr.sync()
        g1, v1, g2, v2, g3, v3 = mk(tr.n_g), mk(tr.nnz), mk(tr.n_g), mk(tr.nnz), mk(tr.n_g), mk(tr.nnz)
        tr.g_jac_dev_peers(zd.data_ptr(), wd.data_ptr(), g1.data_ptr(), v1.data_ptr(), [g2.data_ptr(), g3.data_ptr()],
                           [v2.data_ptr(), v3.data_ptr()])
        tr.sync()
        torch.cuda.synchronize()
        for a, b in ((g1, g0), (g2, g0), (g3, g0), (v1, v0), (v2, v0), (v3, v0)):
            assert torch.equal(a, b)
        if seg is not None:  # rows of the other shards were not written
            assert (g0 == -7.0).any() and (v0 == -7.0).any()


@pytest.mark.parametrize("problem,K,po,scheme", [("van_der_pol", 5, [3, 6, 4, 9, 2], "LGR"), ("synthetic_6_3", 4, 15, "LGL"),
                                                 ("two_phase_schwartz", 3, 7, "CGL"), ("kitchen_sink", 4, [4, 3, 5, 4], "LGR"),
                                                 ("hyper_sensitive", 2, [40, 12], "LGR")])
def test_interpolation_and_residuals_match_oracle(libmpx, problem, K, po, scheme):
    """interpolation of a solution and the dynamics residual at arbitrary per-segment points
    (mpx_eval_residuals) against the oracle's restatement of mpopt.py:1428-1542, ragged and empty segments included."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle import residual as R
    from oracle.nlp import OracleNLP

    tr = Transcription(REGISTRY[problem](), K, po, scheme)
    ora = OracleNLP(REGISTRY[problem](), K, po, scheme)
    z, p = random_point(ora, dirichlet=True)
    rng = np.random.default_rng(2)
    for ph in range(ora.P):
        grids = [R.residual_grid_taus(ora, ph, "mid-points", p), R.residual_grid_taus(ora, ph, "fixed", p),
                 [np.sort(rng.uniform(ora.tau0, ora.tau1, int(rng.integers(0, 6)))) for _ in range(K)],
                 [np.array([ora.tau0, ora.tau1]) if k % 2 == 0 else np.array([]) for k in range(K)]]
        for taus in grids:
            got = tr.residuals(z, p, ph, taus)
            Xi, Ui, ti, DXi, DUi = R.interpolate_phase(ora, z, p, ph, taus)
            _, res, _, n = R.dynamics_residuals_phase(ora, z, p, ph, taus)
            assert got["counts"] == n
            assert_close(got["xi"], Xi, "Xi")
            assert_close(got["ui"], Ui, "Ui")
            assert_close(got["ti"], ti, "ti")
            assert_close(got["dxi"], DXi, "DXi")
            assert_close(got["dui"], DUi, "DUi")
            assert_close(got["res"], res, "residual")


@pytest.mark.parametrize("problem,K,po,scheme", [("moon_lander", 6, 4, "LGR"), ("kitchen_sink", 5, [3, 2, 4, 5, 3], "LGL"),
                                                 ("two_phase_schwartz", 4, 6, "LGR"), ("van_der_pol", 3, [2, 5, 3], "CGL"),
                                                 ("robot_arm", 7, 5, "LGR"), ("synthetic_6_3", 33, 15, "LGR"),
                                                 ("hyper_sensitive", 3, 9, "LGL"), ("generic_two_phase", 3, [2, 5, 3], "CGL"),
                                                 ("hyper_sensitive", 2, [40, 7], "LGR"),
                                                 ("delta3_launch_vehicle", 2, [5, 4], "LGR"),
                                                 ("falcon9_launcher", 3, 4, "LGR")])
def test_lagrangian_hessian_matches_oracle(libmpx, problem, K, po, scheme):
    """nlp_hess_l(x, p, lam_f, lam_g) -- lower triangle, CSR -- against the oracle's second-order
    dual-number restatement: pattern bit-exact, values to 1e-10."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.hessian import hess_l
    from oracle.nlp import OracleNLP

    tr = Transcription(REGISTRY[problem](), K, po, scheme)
    ora = OracleNLP(REGISTRY[problem](), K, po, scheme)
    z, p = random_point(ora, dirichlet=True)
    if problem in ("delta3_launch_vehicle", "falcon9_launcher"):  # near the ascent trajectory (orbital-element terminal rows: acos, 1/|e|)
        z = ora.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(7).standard_normal(ora.n_z))
    rng = np.random.default_rng(4)
    lam, sig = rng.uniform(-1, 1, ora.n_g), 0.6
    H = hess_l(ora, z, p, sig, lam)
    rp, ci = tr.hess_structure()
    assert np.array_equal(rp, H.indptr) and np.array_equal(ci, H.indices), "Hessian pattern differs from the oracle"
    assert_close(tr.hess_l_values(z, p, sig, lam), H.data, "hess_l values")
    # a second point and other multipliers through the same plan (positions are reused, nothing stale is left behind)
    z2, lam2 = z + 1e-2, rng.uniform(-2, 2, ora.n_g)
    assert_close(tr.hess_l_values(z2, p, 1.0, lam2), hess_l(ora, z2, p, 1.0, lam2).data, "hess_l values, second point")


@pytest.mark.parametrize("problem,K,po,scheme", [("van_der_pol", 5, [3, 6, 4, 9, 2], "LGR"), ("kitchen_sink", 3, [5, 7, 4], "LGL"),
                                                 ("synthetic_6_3", 3, 15, "CGL")])
def test_second_derivatives_match_oracle(libmpx, problem, K, po, scheme):
    """mpx_eval_second_derivatives against the oracle's composite D2 (mpopt.py:1285-1358), points on and off the nodes."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle import residual as R
    from oracle.nlp import OracleNLP

    tr = Transcription(REGISTRY[problem](), K, po, scheme)
    ora = OracleNLP(REGISTRY[problem](), K, po, scheme)
    z, p = random_point(ora, dirichlet=True)
    rng = np.random.default_rng(8)
    taus = [np.sort(rng.uniform(-1, 1, int(rng.integers(0, 6)))) for _ in range(ora.K)]
    taus[0] = np.array([-1.0, ora.tab.roots[ora.po[0]][1], 1.0])  # segment ends and a collocation node
    for ph in range(ora.P):
        ti, ddx, ddu = R.second_derivatives_phase(ora, z, p, ph, taus)
        out = tr.second_derivatives(z, p, ph, taus)
        assert out["counts"] == [len(t) for t in taus]
        assert_close(out["ti"], ti, "ti")
        scale = max(1.0, np.abs(ddx).max(), np.abs(ddu).max())
        assert np.abs(out["ddxi"] - ddx).max() <= 1e-9 * scale and np.abs(out["ddui"] - ddu).max() <= 1e-9 * scale


@pytest.mark.parametrize("problem,K,po,scheme", [("van_der_pol", 5, [3, 6, 4, 9, 2], "LGR"), ("kitchen_sink", 3, [5, 7, 4], "LGL"),
                                                 ("synthetic_6_3", 3, 15, "CGL"), ("two_phase_schwartz", 4, 5, "LGR")])
def test_state_residuals_match_oracle(libmpx, problem, K, po, scheme):
    """mpx_eval_state_residuals (mpopt.py:989-1076) against the oracle: segments with 0 .. 8 target points, the
    reference's "spectral" grid, points on the segment ends."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle import residual as R
    from oracle.nlp import OracleNLP

    tr = Transcription(REGISTRY[problem](), K, po, scheme)
    ora = OracleNLP(REGISTRY[problem](), K, po, scheme)
    z, p = random_point(ora, dirichlet=True)
    rng = np.random.default_rng(9)
    custom = [np.sort(rng.uniform(-0.98, 1, int(rng.integers(0, 9)))) for _ in range(ora.K)]
    custom[0] = np.array([-0.6, 0.1, 1.0])
    for taus in (custom, R.residual_grid_taus(ora, 0, "spectral"), R.residual_grid_taus(ora, 0, "mid-points")):
        for ph in range(ora.P):
            xint, res, ti = R.states_from_dynamics_phase(ora, z, p, ph, taus)
            out = tr.state_residuals(z, p, ph, taus)
            assert out["counts"] == [len(t) for t in taus]
            assert_close(out["ti"], ti, "ti")
            assert_close(out["xint"], xint, "xint", 1e-9)
            scale = max(1.0, np.abs(xint).max())
            assert np.abs(out["res_x"] - res).max() <= 1e-9 * scale


@pytest.mark.parametrize("name,K,po,scheme", [("alp_rider", 10, 5, "LGR"), ("mine_opt", 1, 30, "LGR"),
                                              ("dae_van_der_pol", 50, 3, "LGR")])
def test_reference_examples_run_time_compiled(libmpx, name, K, po, scheme):
    """The remaining problems of the reference's tests/test_examples.py:38-50 at the sizes those modules build
    (alpr01 10 x 5, mineopt 1 x 30, vdp 50 x 3): not registered, so their functors go through NVRTC; g, jac_g, f,
    grad_f and hess_l against the oracle."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import EXAMPLES
    from oracle.hessian import hess_l
    from oracle.nlp import OracleNLP

    tr = Transcription(EXAMPLES[name](), K, po, scheme, drop_exact_zeros=False)
    ora = OracleNLP(EXAMPLES[name](), K, po, scheme, drop_exact_zeros=False)
    assert tr.program_origin.startswith("nvrtc:")
    z, p = random_point(ora, dirichlet=True)
    if name == "mine_opt":
        z = np.abs(z) + 0.5  # the cost divides by the state
    J = ora.jac_g(z, p)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(z, p, g_out=g), J.data, "jac_g values")
    assert_close(g, ora.g(z, p), "g")
    assert abs(tr.f(z, p) - ora.f(z, p)) <= 1e-10 * max(1.0, abs(ora.f(z, p)))
    assert_close(tr.grad_f(z, p), ora.grad_f(z, p), "grad_f")
    lam = np.random.default_rng(3).uniform(-1, 1, ora.n_g)
    H = hess_l(ora, z, p, 0.7, lam)
    hrp, hci = tr.hess_structure()
    assert np.array_equal(hrp, H.indptr) and np.array_equal(hci, H.indices)
    assert_close(tr.hess_l_values(z, p, 0.7, lam), H.data, "hess_l values")


# ----------------------------------------------------------------------------- host hop: registered buffers, dynamic fetch
@pytest.mark.gpu
@pytest.mark.parametrize("make,K,po,scheme", [
    ("synthetic_6_3", 6, 5, "LGR"), ("kitchen_sink", 4, [3, 5, 4, 3], "LGR"), ("two_phase_schwartz", 3, 4, "LGL"),
    ("moon_lander", 5, 4, "LGL"), ("delta3_launch_vehicle", 1, 5, "LGR")])
def test_dynamic_fetch_equals_full_fetch(libmpx, make, K, po, scheme):
    """mpx_eval_jac_g_dynamic into a registered buffer: first call = full fetch, later calls rewrite only the z- / p-
    dependent entries -- and the buffer still equals a full evaluation at every new point.  The positions reported as
    dynamic are exactly the entries that ever change."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    ocp = REGISTRY[make]()
    tr = Transcription(ocp, K, po, scheme)
    ora = OracleNLP(ocp, K, po, scheme)
    z, w = random_point(ora, dirichlet=True)
    vals, g = np.full(tr.nnz, np.nan), np.empty(tr.n_g)
    tr.host_register(vals)
    try:
        pos = tr.dynamic_positions()
        assert (np.diff(pos) > 0).all() and 0 < len(pos) < tr.nnz
        full0 = tr.jac_g_values(z, w)
        tr.jac_g_values_dynamic(z, w, out=vals, g_out=g)          # primes the buffer (full fetch)
        assert np.array_equal(vals, full0)
        changed = np.zeros(tr.nnz, bool)
        rng = np.random.default_rng(5)
        for it in range(3):
            z2 = z + 0.05 * rng.standard_normal(z.size)
            w2 = rng.dirichlet(np.ones(ora.K), size=ora.P).reshape(-1)
            full = tr.jac_g_values(z2, w2)
            before = vals.copy()
            tr.jac_g_values_dynamic(z2, w2, out=vals, g_out=g)
            assert np.array_equal(vals, full), f"dynamic fetch differs from the full fetch at point {it}"
            assert_close(g, ora.g(z2, w2), "g of the dynamic fetch")
            changed |= before != vals
            packed = tr.jac_g_packed(z2, w2)
            assert np.array_equal(packed, full[pos])
        const = np.ones(tr.nnz, bool)
        const[pos] = False
        assert not changed[const].any(), "an entry outside the dynamic set changed"
        assert_close(vals, ora.jac_g(z2, w2).data, "dynamic fetch vs oracle")
    finally:
        tr.host_unregister(vals)


@pytest.mark.gpu
def test_dynamic_fetch_tracks_the_buffer_it_primed(libmpx):
    """Unregistered (pageable) buffers work too: the first call on a buffer writes everything, later calls on the SAME
    buffer only the dynamic entries; handing in ANOTHER buffer is noticed and answered with a full fetch."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import moon_lander

    tr = Transcription(moon_lander(), 4, 3, "LGR")
    rng = np.random.default_rng(1)
    z = rng.uniform(-1, 1, tr.n_z)
    z[-2:] = [0.0, 3.0]
    a, b = np.full(tr.nnz, np.nan), np.full(tr.nnz, np.nan)
    tr.jac_g_values_dynamic(z, out=a)
    assert np.array_equal(a, tr.jac_g_values(z))
    z2 = z + 0.1
    tr.jac_g_values_dynamic(z2, out=a)              # same buffer: dynamic entries only
    assert np.array_equal(a, tr.jac_g_values(z2))
    tr.jac_g_values_dynamic(z2, out=b)              # another buffer: everything
    assert np.array_equal(b, a)
    pos = tr.dynamic_positions()
    b[pos] = np.nan
    tr.jac_g_values_dynamic(z, out=b)               # b is primed now: its dynamic entries come back
    assert np.array_equal(b, tr.jac_g_values(z))


@pytest.mark.gpu
@pytest.mark.parametrize("make,K,po,world", [("kitchen_sink", 6, [3, 5, 4, 3, 2, 4], 3), ("moon_lander", 8, 3, 2),
                                             ("synthetic_6_3", 6, 4, 3)])
def test_sharded_objective_partials_sum_to_the_full_objective(libmpx, make, K, po, world):
    """f + grad_f of shard plans (the objective side of the multi-GPU exchange): partial J and partial d/dt0, d/dtf,
    d/da summed over the shards, node entries taken from their owners, Mayer x0 entries from the last shard --
    exactly what mpopt_b200.shard.ObjectiveGatherer does with one all-reduce and one all-gather -- equal the
    evaluation of the whole NLP."""
    import torch

    from mpopt_b200 import shard as sh
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    ocp = REGISTRY[make]()
    ora = OracleNLP(ocp, K, po, "LGR")
    z, w = random_point(ora, dirichlet=True)
    pol = [po] * K if isinstance(po, int) else list(po)
    part = sh.partition(pol, world)
    dev = torch.device("cuda", 0)
    zd, wd = torch.from_numpy(z).to(dev), torch.from_numpy(w).to(dev)
    fs, grads, og = [], [], None
    for r, seg in enumerate(part):
        tr = Transcription(ocp, K, po, "LGR", segments=seg)
        og = og or sh.ObjectiveGatherer(tr.layout, part, None, 0)
        f = torch.zeros(1, dtype=torch.float64, device=dev)
        grad = torch.full((tr.n_z,), float("nan"), dtype=torch.float64, device=dev)
        tr.f_grad_dev(zd.data_ptr(), wd.data_ptr(), f.data_ptr(), grad.data_ptr())
        torch.cuda.synchronize()
        fs.append(float(f[0])), grads.append(grad.cpu().numpy())
    glob = np.asarray(og.glob)
    per, nx = len(glob) // ora.P, ora.nx
    x0 = np.zeros(len(glob), bool)
    for ph in range(ora.P):
        x0[ph * per + per - nx:(ph + 1) * per] = True
    total = np.full(ora.n_z, np.nan)
    for r in range(world):
        for off, cnt in og.runs[r]:
            total[off:off + cnt] = grads[r][off:off + cnt]
    small = np.zeros(len(glob))
    for r, (kb, ke) in enumerate(part):
        v = grads[r][glob].copy()
        if not (kb == 0 or ke == K):
            v[x0] = 0.0
        small += v
    total[glob] = small
    assert abs(sum(fs) - ora.f(z, w)) <= 1e-10 * max(1.0, abs(ora.f(z, w)))
    assert_close(total, ora.grad_f(z, w), "gathered grad_f")


@pytest.mark.gpu
def test_residuals_accept_a_2d_array_of_points(libmpx):
    """One [K, m] array of local abscissae (same number of points in every segment) gives what the reference's
    per-segment lists give."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import moon_lander

    K = 7
    tr = Transcription(moon_lander(), K, 4, "LGL")
    rng = np.random.default_rng(3)
    z = rng.uniform(-1, 1, tr.n_z)
    z[-2:] = [0.0, 3.0]
    taus = rng.uniform(-1, 1, (K, 5))
    a = tr.residuals(z, None, 0, [taus[k] for k in range(K)])
    b = tr.residuals(z, None, 0, taus)
    assert a["counts"] == b["counts"]
    for key in ("xi", "ui", "ti", "dxi", "dui", "res"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.gpu
def test_hessian_row_runs_equal_scattered_stores_at_full_size(libmpx, monkeypatch):
    """hess_l at the headline size (65 537 nodes, 2.58 M entries): the default path (affine positions by value, node rows
    staged per warp and written as contiguous runs) writes the same bits as the entry-by-entry stores, the values are
    linear in the multipliers, and a second call through the same plan leaves nothing stale behind."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import synthetic_6_3

    K, po = 4096, 15
    tr = Transcription(synthetic_6_3(), K, po, "LGR")
    monkeypatch.setenv("MPX_HESS_ROWS", "0")
    tr0 = Transcription(synthetic_6_3(), K, po, "LGR")
    rp0, ci0 = tr0.hess_structure()  # the switch is read when the Hessian plan is built
    monkeypatch.delenv("MPX_HESS_ROWS")
    rp, ci = tr.hess_structure()
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and len(ci) == 2580522
    rng = np.random.default_rng(5)
    z = rng.uniform(-1, 1, tr.n_z)
    z[-2:] = [0.0, 1.0]
    p = rng.dirichlet(np.ones(K))
    l1, l2 = rng.uniform(-1, 1, tr.n_g), rng.uniform(-1, 1, tr.n_g)
    h1 = tr.hess_l_values(z, p, 0.7, l1).copy()
    assert np.array_equal(h1, tr0.hess_l_values(z, p, 0.7, l1)), "row runs and scattered stores differ"
    h2 = tr.hess_l_values(z, p, -0.3, l2).copy()
    h12 = tr.hess_l_values(z, p, 2.0 * 0.7 - 0.5 * -0.3, 2.0 * l1 - 0.5 * l2)
    assert_close(h12, 2.0 * h1 - 0.5 * h2, "linearity in (lam_f, lam_g)", 1e-11)
    assert np.array_equal(tr.hess_l_values(z, p, 0.7, l1), h1), "a second evaluation through the same plan differs"
    assert np.all(np.isfinite(h1)) and np.count_nonzero(h1) > 0.5 * len(h1)
