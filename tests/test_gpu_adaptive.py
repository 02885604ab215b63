"""GPU parity of the widths-as-variables NLP (``mp.mpopt_adaptive``) against oracle/adaptive.py, and
the reference's adaptive tests (tests/test_mpopt.py:431-483, :498-546) re-read against ``mpopt_b200.mp``: the
h-adaptive outer loop re-solves one device plan with new width parameters, the adaptive class solves for the widths."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

CASES = [
    # name, problem, K, poly_orders, scheme, mid_residuals
    ("moon_3x3", "moon_lander", 3, 3, "LGR", True),          # tests/test_mpopt.py:258-259
    ("moon_nores", "moon_lander", 3, 3, "LGR", False),       # tests/test_mpopt.py:474
    ("hyper_5x15", "hyper_sensitive", 5, 15, "LGR", True),   # tests/test_mpopt.py:276-277
    ("hyper_3x30", "hyper_sensitive", 3, 30, "LGR", True),   # examples/singlephase/hyper_sensitive.py:82
    ("vdp_mixed", "van_der_pol", 6, [3, 7, 2, 5, 4, 6], "CGL", True),
    ("schwartz", "two_phase_schwartz", 4, 6, "LGL", True),
    ("syn63", "synthetic_6_3", 12, 15, "LGR", True),
    ("robot", "robot_arm", 5, 4, "LGR", True),
    ("sink_time_dependent", "kitchen_sink", 5, [3, 4, 6, 2, 5], "LGR", True),
    ("sink_p1", "kitchen_sink", 3, 1, "LGL", True),
    ("delta3", "delta3_launch_vehicle", 3, [4, 6, 5], "LGR", True),
    ("falcon9", "falcon9_launcher", 2, [5, 3], "LGR", True),
]


def _point(n, problem, seed=20261017):
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, n.n_z)
    if problem == "robot_arm":
        z = np.abs(z) + 0.5
    if problem in ("delta3_launch_vehicle", "falcon9_launcher"):
        return n.initialize_solution() * (1.0 + 0.01 * rng.standard_normal(n.n_z))
    for ph in range(n.P):
        z[n.colT0(ph)] = 0.3 + 0.25 * ph
        z[n.colTF(ph)] = 2.0 + 1.5 * ph
        for m in range(n.na):
            z[n.colA(ph, m)] = rng.uniform(0.2, 1.2)
        z[n.colW(ph, np.arange(n.K))] = rng.dirichlet(np.ones(n.K)) * 0.8 + 0.2 / n.K
    return z


@pytest.mark.parametrize("name,problem,K,po,scheme,mid", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("drop", [False, True], ids=["structural", "folded"])
def test_adaptive_nlp_matches_oracle(libmpx, name, problem, K, po, scheme, mid, drop):
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.adaptive import OracleAdaptiveNLP

    ocp = REGISTRY[problem]()
    tr = Transcription(ocp, K, po, scheme, adaptive=True, mid_residuals=mid, drop_exact_zeros=drop)
    tabs = None
    if drop:  # exact-zero folding is decided on the device's own tables (checked against the oracle's first)
        plain = OracleAdaptiveNLP(ocp, K, po, scheme, drop_exact_zeros=False, mid_residuals=mid)
        tabs = {}
        for d in sorted(set(tr.poly_orders)):
            r, D, w, Cm = tr.tables(d)
            assert_close(D, plain.tab.D[d], f"D[{d}]", 1e-11)
            assert_close(Cm, plain.tab.Cmid[d], f"Cmid[{d}]", 1e-12)
            tabs[d] = (r, D, w, Cm)
    ora = OracleAdaptiveNLP(ocp, K, po, scheme, drop_exact_zeros=drop, mid_residuals=mid, tables=tabs)
    assert (tr.n_z, tr.n_p, tr.n_g) == (ora.n_z, 0, ora.n_g)
    assert tr.program_origin.split(";")[-1] in ("adaptive", "adaptive/in-place")
    z = _point(ora, problem)
    J = ora.jac_g(z)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr), "rowptr differs from the oracle"
    assert np.array_equal(ci, J.indices), "colind differs from the oracle"
    g = np.empty(tr.n_g)
    vals = tr.jac_g_values(z, g_out=g)
    assert_close(g, ora.g(z), "g")
    assert_close(tr.g(z), ora.g(z), "g (g-only launch)")
    assert_close(vals, J.data, "jac_g values")
    assert abs(tr.f(z) - ora.f(z)) <= 1e-10 * max(1.0, abs(ora.f(z)))
    assert_close(tr.grad_f(z), ora.grad_f(z), "grad_f")
    for a, b in zip(tr.bounds(), ora.bounds()):
        assert np.array_equal(a, b)
    assert np.allclose(tr.initial_guess(), ora.initialize_solution(), rtol=0, atol=1e-15)
    # a second point: nothing cached between evaluations
    z2 = z + 1e-3 * np.random.default_rng(1).standard_normal(tr.n_z)
    assert_close(tr.jac_g_values(z2), ora.jac_g(z2).data, "jac_g values, second point")


def test_adaptive_nlp_runtime_compiled_program(libmpx):
    """An unregistered problem goes through NVRTC: the adaptive kernels are part of the run-time module."""
    from mpopt_b200 import ca
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.ocp import OCP
    from oracle.adaptive import OracleAdaptiveNLP

    ocp = OCP(n_states=2, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [x[1] * ca.cos(0.37 * t), u[0] - 0.21 * x[0] * x[1]]
    ocp.running_costs[0] = lambda x, u, t: u[0] * u[0] + 0.13 * t * x[0]
    ocp.lbu[0], ocp.ubu[0] = -1, 1
    ocp.validate()
    tr = Transcription(ocp, 4, [3, 5, 2, 4], "LGR", adaptive=True, drop_exact_zeros=False)
    ora = OracleAdaptiveNLP(ocp, 4, [3, 5, 2, 4], "LGR", drop_exact_zeros=False)
    assert tr.program_origin.startswith("nvrtc:")
    z = _point(ora, "x")
    J = ora.jac_g(z)
    rp, ci = tr.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    assert_close(tr.jac_g_values(z), J.data, "jac_g values")
    assert_close(tr.g(z), ora.g(z), "g")
    assert_close(tr.grad_f(z), ora.grad_f(z), "grad_f")


def test_adaptive_plan_refuses_what_it_cannot_do(libmpx):
    from mpopt_b200 import _lib
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import moon_lander

    with pytest.raises(_lib.MpxError):
        Transcription(moon_lander(), 4, 3, "LGR", adaptive=True, segments=(0, 2))
    from mpopt_b200.problems import kitchen_sink

    tr = Transcription(kitchen_sink(), 4, [3, 5, 4, 3], "LGR", adaptive=True)  # explicit time dependence
    with pytest.raises(_lib.MpxError):
        tr.hess_structure()


HESS_CASES = [("moon_lander", 3, 3, "LGR", True), ("moon_lander", 4, [4, 2, 3, 5], "LGL", True),
              ("hyper_sensitive", 3, [4, 2, 3], "LGR", True), ("hyper_sensitive", 5, 15, "CGL", True),
              ("synthetic_6_3", 2, 3, "LGR", True), ("synthetic_6_3", 5, 6, "LGR", False),
              ("van_der_pol", 3, 5, "LGR", True), ("two_phase_schwartz", 2, 4, "LGR", True),
              ("robot_arm", 2, 4, "LGR", True),
              # uniform degrees with a degree-specialised instance of mpx_adapt_hess_kernel (problems.py::AOT_DEGREES)
              ("synthetic_6_3", 3, 15, "LGR", True), ("moon_lander", 4, 15, "LGL", True), ("two_phase_schwartz", 3, 10, "CGL", True),
              # degrees above 15: the dense blocks leave the tensor-core path for the per-warp product tiles
              ("hyper_sensitive", 2, 20, "LGR", True), ("van_der_pol", 3, [18, 4, 17], "LGL", True),
              ("synthetic_6_3", 2, [16, 15], "CGL", True)]


@pytest.mark.parametrize("problem,K,po,scheme,mid", HESS_CASES, ids=[f"{c[0]}-{c[1]}-{c[3]}-{'res' if c[4] else 'nores'}" for c in HESS_CASES])
def test_adaptive_hessian_matches_oracle(libmpx, problem, K, po, scheme, mid):
    """nlp_hess_l of the widths-as-variables NLP (what ca.nlpsol derives for mpopt_adaptive, mpopt.py:757 with the NLP
    of :3174-3205): lower triangle, pattern bit-exact, values to 1e-10 against the second-order-dual oracle (which is
    itself checked by finite differences, tests/test_oracle_adaptive.py).  Covers the bilinear h_k = (tf - t0) w_k /
    delta terms and the dense per-segment blocks of the mid-point residual rows."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.adaptive import OracleAdaptiveNLP
    from oracle.hessian import hess_l

    ocp = REGISTRY[problem]()
    tr = Transcription(ocp, K, po, scheme, adaptive=True, mid_residuals=mid)
    ora = OracleAdaptiveNLP(ocp, K, po, scheme, mid_residuals=mid)
    z = _point(ora, problem)
    rng = np.random.default_rng(9)
    lam = rng.uniform(-1, 1, tr.n_g)
    import scipy.sparse as sp

    H = hess_l(ora, z, None, 0.7, lam)
    rp, ci = tr.hess_structure()
    assert np.array_equal(rp, H.indptr) and np.array_equal(ci, H.indices), "Hessian pattern differs from the oracle"
    assert_close(tr.hess_l_values(z, None, 0.7, lam), H.data, "hess_l values")
    # constraints only (lam_f = 0; the oracle's pattern then loses the entries that only the objective has)
    H0 = hess_l(ora, z, None, 0.0, lam)
    G0 = sp.csr_matrix((tr.hess_l_values(z, None, 0.0, lam), ci, rp), shape=H0.shape)
    diff = abs(G0 - H0)
    assert diff.max() <= 1e-10 * max(1.0, abs(H0).max())
    # evaluating twice gives the same bits (the += accumulation is ordered)
    assert np.array_equal(tr.hess_l_values(z, None, 0.7, lam), tr.hess_l_values(z, None, 0.7, lam))
    # the first evaluation probed (on a NaN-filled buffer) whether every entry of the pattern has a writer; none of
    # these problems needs the zero fill, and nothing of an earlier evaluation survives in the device buffer
    assert tr.hess_zero_fill == 0
    z2 = z * 0.9 + 0.01
    lam2 = np.random.default_rng(10).uniform(-1, 1, tr.n_g)
    assert_close(tr.hess_l_values(z2, None, 0.3, lam2), hess_l(ora, z2, None, 0.3, lam2).data, "hess_l values, second point")


@pytest.mark.parametrize("po", [4, [3, 5, 2, 4, 6], 17])
def test_adaptive_hessian_runtime_compiled_program(libmpx, po):
    """The Hessian kernels of an unregistered problem come out of NVRTC (generic-degree instance, the run-time launcher
    sizes the shared memory itself): same checks as the ahead-of-time programs."""
    from mpopt_b200 import ca
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.ocp import OCP
    from oracle.adaptive import OracleAdaptiveNLP
    from oracle.hessian import hess_l

    ocp = OCP(n_states=2, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [x[1] * ca.cos(0.37 * x[0]), u[0] - 0.21 * x[0] * x[1] * u[0]]
    ocp.running_costs[0] = lambda x, u, t: u[0] * u[0] + 0.13 * x[0] * x[0] * x[1]
    ocp.terminal_costs[0] = lambda xf, tf, x0, t0: xf[0] * xf[1] + tf * xf[0]
    ocp.lbu[0], ocp.ubu[0] = -1, 1
    ocp.validate()
    K = 5 if isinstance(po, list) else 3
    tr = Transcription(ocp, K, po, "LGL", adaptive=True)
    ora = OracleAdaptiveNLP(ocp, K, po, "LGL")
    assert tr.program_origin.startswith("nvrtc:")
    z = _point(ora, "x")
    lam = np.random.default_rng(3).uniform(-1, 1, tr.n_g)
    H = hess_l(ora, z, None, 0.6, lam)
    rp, ci = tr.hess_structure()
    assert np.array_equal(rp, H.indptr) and np.array_equal(ci, H.indices)
    assert_close(tr.hess_l_values(z, None, 0.6, lam), H.data, "hess_l values")
    assert tr.hess_zero_fill == 0
    z2 = z * 0.8 - 0.02
    assert_close(tr.hess_l_values(z2, None, 0.6, lam), hess_l(ora, z2, None, 0.6, lam).data, "hess_l values, second point")


@pytest.mark.parametrize("problem,K,po,scheme", [("synthetic_6_3", 7, 15, "LGR"), ("moon_lander", 6, [4, 2, 3, 5, 4, 3], "LGL"),
                                                 ("two_phase_schwartz", 4, 10, "CGL")])
def test_adaptive_hessian_launch_variants_agree_bit_for_bit(libmpx, monkeypatch, problem, K, po, scheme):
    """The shipped path (ONE persistent launch with flags at the shared nodes, node-local block diagonals staged and
    written with their runs) against the plainer ones it replaced: one launch per segment parity (MPX_AHESS_PERSIST=0),
    node-local entries stored by mpx_hess_kernel and added to afterwards (MPX_AHESS_STAGE=0), the zero-filled buffer
    (MPX_AHESS_ZERO=1).  Every entry is the same sum in the same order: the bits must agree."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from oracle.adaptive import OracleAdaptiveNLP

    ocp = REGISTRY[problem]()
    ora = OracleAdaptiveNLP(ocp, K, po, scheme)
    z = _point(ora, problem)
    lam = np.random.default_rng(5).uniform(-1, 1, ora.n_g)
    vals = {}
    for name, env in (("shipped", {}), ("two launches", {"MPX_AHESS_PERSIST": "0"}), ("no staging", {"MPX_AHESS_STAGE": "0"}),
                      ("plain", {"MPX_AHESS_PERSIST": "0", "MPX_AHESS_STAGE": "0", "MPX_AHESS_ZERO": "1"})):
        for k_ in ("MPX_AHESS_PERSIST", "MPX_AHESS_STAGE", "MPX_AHESS_ZERO"):
            monkeypatch.delenv(k_, raising=False)
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        tr = Transcription(ocp, K, po, scheme, adaptive=True)
        tr.hess_structure()
        vals[name] = [tr.hess_l_values(z, None, 0.7, lam).copy() for _ in range(3)]  # graph-free repeats: flags / epochs advance
        assert tr.hess_zero_fill == (1 if env.get("MPX_AHESS_ZERO") else 0)
        del tr
    for name, reps in vals.items():
        for r in reps:
            assert np.array_equal(r, vals["shipped"][0]), name


# ---------------------------------------------------------------------------- the reference's adaptive tests
@pytest.fixture(scope="module")
def mp(libmpx):
    from mpopt_b200 import mp as _mp

    _mp.mpopt._MUTE_ = True
    return _mp


def _check_post(mpo, sol):
    for key in ("x", "f"):
        assert key in sol
    post = mpo.process_results(sol, plot=False)
    x, u, t, _ = post.get_data()
    xi, ui, ti, _ = post.get_data(interpolate=True)
    assert x.shape[0] == u.shape[0] == t.shape[0]
    assert xi.shape[0] == ui.shape[0] == ti.shape[0]
    return post


@pytest.mark.parametrize("grid,max_iter,options", [
    ("fixed", 3, {}),                                                            # tests/test_mpopt.py:431-440
    ("mid-points", 2, {"method": "residual", "sub_method": "equal_area"}),      # :443-455
    ("spectral", 10, {"method": "control_slope", "sub_method": ""}),            # :458-470
    ("fixed", 3, {"method": "residual", "sub_method": "merge_split"}),
])
def test_moon_lander_h_adaptive_solve(mp, grid, max_iter, options):
    from mpopt_b200.problems import moon_lander

    mpo = mp.mpopt_h_adaptive(moon_lander(), 10, 4)
    mpo.grid_type[0] = grid
    plan = mpo.transcription
    sol = mpo.solve(max_iter=max_iter, mpopt_options=dict(options))
    assert mpo.transcription is plan, "the refinement loop must reuse the device plan (widths are parameters)"
    assert 1 <= mpo.iter_count <= max_iter
    sw = np.asarray(mpo._nlp_sw_params, float)
    assert sw.shape == (10,) and abs(sw.sum() - 1) < 1e-9 and (sw > 0).all()
    assert abs(sol["f"] - 8.2477) < 5e-2   # docs/source/notebooks/moon_lander.ipynb:185
    _check_post(mpo, sol)


def test_h_adaptive_reduces_the_residual(mp):
    """The point of the loop: moving the segment boundaries towards the bang-bang switch lowers the max residual."""
    from mpopt_b200.problems import moon_lander

    mpo = mp.mpopt_h_adaptive(moon_lander(), 10, 4)
    mpo.solve(max_iter=4, mpopt_options={"method": "residual", "sub_method": "merge_split"})
    errs = [v for v in mpo.iter_info.values() if v is not None]
    assert len(errs) >= 2 and errs[-1] < errs[0]


def test_hyper_sensitive_h_adaptive_solve(mp):
    """tests/test_mpopt.py:268-269, :498-507 at a size SciPy's solver handles in seconds."""
    from mpopt_b200.problems import hyper_sensitive

    mpo = mp.mpopt_h_adaptive(hyper_sensitive(), 6, 8)
    sol = mpo.solve(max_iter=2, mpopt_options={"method": "residual", "sub_method": "merge_split"})
    _check_post(mpo, sol)


def test_moon_lander_mpopt_adaptive_solve(mp):
    """tests/test_mpopt.py:258-259, :473-483."""
    from mpopt_b200.problems import moon_lander

    mpo = mp.mpopt_adaptive(moon_lander(), 3, 3)
    mpo.mid_residuals = False
    sol = mpo.solve()
    post = _check_post(mpo, sol)
    sw = mpo._nlp_sw_params
    assert sw.shape == (3,) and abs(sw.sum() - 1) < 1e-6 and (sw >= 1e-4 - 1e-9).all()
    zmin, zmax, gmin, gmax = mpo.transcription.bounds()
    assert (sol["g"] >= gmin - 1e-5).all() and (sol["g"] <= gmax + 1e-5).all()
    assert abs(sol["f"] - 8.2477) < 0.2
    x, u, t, _ = post.get_data()
    assert abs(x[0, 0] - 10.0) < 1e-9 and abs(x[-1, 0]) < 1e-5
    # the trajectories of ANOTHER solution vector are laid out on that vector's own widths, not on those of the last
    # solve (the reference evaluates the time grid from Z: mpopt.py:3248-3273)
    L = mpo.transcription.layout
    z2 = np.array(sol["x"], dtype=float).reshape(-1).copy()
    w2 = np.array([0.5, 0.2, 0.3])
    z2[L.colW(0, 0): L.colW(0, 0) + 3] = w2
    post2 = mpo.process_results({**sol, "x": z2}, plot=False)
    _, _, t2, _ = post2.get_data()
    t0, tf = z2[L.colT0(0)], z2[L.colTF(0)]
    ends = t2.ravel()[[3, 6, 9]]
    assert np.allclose(ends, t0 + (tf - t0) * np.cumsum(w2), rtol=0, atol=1e-12)
    assert np.allclose(mpo._nlp_sw_params, sw)  # the last solve's widths are untouched


def test_hyper_sensitive_mpopt_adaptive_structure(mp):
    """tests/test_mpopt.py:276-277: mpopt_adaptive(hyper_sensitive, 5, 15) -- sizes of the NLP the reference builds."""
    from mpopt_b200.problems import hyper_sensitive

    mpo = mp.mpopt_adaptive(hyper_sensitive(), 5, 15)
    nlp, bounds = mpo.create_nlp()
    N = 76
    assert nlp["x"] == 2 * N + 2 + 5 and nlp["p"] == 0
    # F N | TC 1 | sum 1 | residuals N-1   (no finite state / control bounds in this problem)
    assert bounds["lbg"].shape[0] == N + 1 + 1 + (N - 1)
    SW, SWmin, SWmax = mpo.get_nlp_constrains_for_segment_widths(0)
    assert len(SW) == 1 + (N - 1) and SWmin[0] == SWmax[0] == 0 and (SWmax[1:] == 1e-3).all()
    assert mpo.initialize_solution().shape[0] == nlp["x"]


def test_h_adaptive_width_update_two_phases(mp):
    """Width update of a two-phase problem from a given point (no solve): per-phase widths stay normalised, the
    residuals come from the GPU kernel, every method returns one width per segment and phase (mpopt.py:2474-2592)."""
    from mpopt_b200.problems import two_phase_schwartz

    mpo = mp.mpopt_h_adaptive(two_phase_schwartz(), 5, 4)
    mpo.create_solver()
    rng = np.random.default_rng(3)
    z = mpo.initialize_solution() + 0.05 * rng.standard_normal(mpo.transcription.n_z)
    sol = {"x": z}
    for options in ({"method": "residual", "sub_method": "equal_area"}, {"method": "residual", "sub_method": "merge_split"},
                    {"method": "control_slope"}):
        w, err = mpo.get_segment_width_parameters(sol, options=options)
        w = np.asarray(w, float)
        assert w.shape == (10,) and err is not None and err > 0
        assert abs(w[:5].sum() - 1) < 1e-9 and abs(w[5:].sum() - 1) < 1e-9 and (w > 0).all()
    ti, res = mpo.get_dynamics_residuals(sol)
    assert len(res) == 2 and len(res[0]) == 5
    assert abs(max(np.abs(r).max() for ph in res for r in ph if r is not None) - err) < 1e-12


@pytest.mark.gpu
def test_width_column_is_written_in_place(libmpx):
    """Dynamics without explicit time dependence: the F rows carry ONE width column, written by the base kernel, so the
    whole Jacobian lands at its CSR positions without a gather pass (2 launches: base kernel + mpx_adapt_kernel).
    Time-dependent dynamics keep the staged path (rows dense in the earlier widths).  Both match the oracle (above)."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import kitchen_sink, moon_lander

    tr = Transcription(moon_lander(), 6, 4, "LGR", adaptive=True)
    assert tr.program_origin.endswith("adaptive/in-place")
    rng = np.random.default_rng(0)
    z = rng.uniform(0.1, 1.0, tr.n_z)
    l0 = tr.launches
    tr.jac_g_values(z)
    assert tr.launches - l0 == 2
    tr_t = Transcription(kitchen_sink(), 4, [3, 5, 4, 3], "LGR", adaptive=True)  # f depends on t
    assert tr_t.program_origin.endswith(";adaptive")


@pytest.mark.gpu
def test_persistent_kernel_equals_one_cta_per_segment(libmpx, monkeypatch):
    """mpx_adapt_kernel at scale (1024 segments of degree 15, and a mixed-degree plan): persistent CTAs taking segments
    off the global counter (tables resident, next segment prefetched) write the same bits as one CTA per segment, for
    g alone and for g + jac_g, and repeatedly through one plan (the counters re-arm themselves)."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import synthetic_6_3, van_der_pol

    for make, K, po in ((synthetic_6_3, 1024, 15), (van_der_pol, 700, [3, 9, 9, 4, 15, 15, 2] * 100)):
        tr = Transcription(make(), K, po, "LGR", adaptive=True)
        monkeypatch.setenv("MPX_QUEUE", "0")
        tr0 = Transcription(make(), K, po, "LGR", adaptive=True)
        monkeypatch.delenv("MPX_QUEUE")
        rng = np.random.default_rng(3)
        z = rng.uniform(-1, 1, tr.n_z)
        L = tr.layout
        z[L.colT0(0)], z[L.colTF(0)] = 0.0, 2.0
        z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
        g, g0 = np.empty(tr.n_g), np.empty(tr.n_g)
        v = tr.jac_g_values(z, None, g_out=g).copy()
        v0 = tr0.jac_g_values(z, None, g_out=g0)
        assert np.array_equal(v, v0) and np.array_equal(g, g0), f"{make.__name__}: persistent and per-segment CTAs differ"
        assert_close(tr.g(z, None), g, "g-only instance", 1e-13)
        for _ in range(3):
            assert np.array_equal(tr.jac_g_values(z, None), v), "a later evaluation through the same plan differs"


@pytest.mark.gpu
def test_adaptive_hessian_is_linear_in_the_multipliers_at_scale(libmpx):
    """Hessian of the widths-as-variables NLP at 512 segments of degree 15 (2.8 M entries; dense per-segment blocks on
    the fp64 tensor cores): H(a l1 + b l2) = a H(l1) + b H(l2), evaluating twice gives the same bits, and the block
    part agrees with the product-tile path used above degree 15 through the small cases of
    test_adaptive_hessian_matches_oracle."""
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import synthetic_6_3

    K = 512
    tr = Transcription(synthetic_6_3(), K, 15, "LGR", adaptive=True)
    rng = np.random.default_rng(9)
    z = rng.uniform(-1, 1, tr.n_z)
    L = tr.layout
    z[L.colT0(0)], z[L.colTF(0)] = 0.0, 1.0
    z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
    l1, l2 = rng.uniform(-1, 1, tr.n_g), rng.uniform(-1, 1, tr.n_g)
    h1 = tr.hess_l_values(z, None, 0.7, l1).copy()
    h2 = tr.hess_l_values(z, None, -0.4, l2).copy()
    h12 = tr.hess_l_values(z, None, 3.0 * 0.7 + 0.25 * -0.4, 3.0 * l1 + 0.25 * l2)
    assert_close(h12, 3.0 * h1 + 0.25 * h2, "linearity in (lam_f, lam_g)", 1e-10)
    assert np.array_equal(tr.hess_l_values(z, None, 0.7, l1), h1), "two evaluations differ (ordering of the sums)"
    assert np.all(np.isfinite(h1))
