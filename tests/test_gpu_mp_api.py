"""The reference's own tests for this path (tests/test_mpopt.py), re-read against ``mpopt_b200.mp``:
collocation known answers, layout / bound length agreements, and end-to-end solves whose NLP callbacks are the
CUDA evaluators (the optimiser is SciPy: IPOPT is not installable here)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mp(libmpx):
    from mpopt_b200 import mp as _mp

    _mp.mpopt._MUTE_ = True
    yield _mp
    _mp.CollocationRoots._TAU_MIN = -1


@pytest.mark.parametrize("scheme", ["LGR", "LGL", "CGL"])
@pytest.mark.parametrize("tau_min", [-1, 0])
def test_collocation_degree_one(mp, scheme, tau_min):
    """tests/test_mpopt.py:927-1086 (the tests set CollocationRoots._TAU_MIN before constructing)."""
    mp.CollocationRoots._TAU_MIN = tau_min
    try:
        col = mp.Collocation([1], scheme)
        assert (np.abs(col.roots[1] - np.array([mp.CollocationRoots._TAU_MIN, mp.CollocationRoots._TAU_MAX])) < 1e-6).all()
        h = col.roots[1][-1] - col.roots[1][0]
        C = col.get_interpolation_matrix(col.roots[1], 1)
        assert C[0, 0] == 1 and C[1, 0] == 0 and C[0, 1] == 0 and C[1, 1] == 1
        D = col.get_diff_matrix(1, order=1)
        assert (np.abs(D - np.array([[-1 / h, 1 / h], [-1 / h, 1 / h]])) < 1e-6).all()
        assert (np.abs(col.get_diff_matrix(1, order=2)) < 1e-6).all()
    finally:
        mp.CollocationRoots._TAU_MIN = -1


def test_collocation_basis_end_points_and_numerical_mode(mp):
    """tests/test_mpopt.py:627-634 and :612-624 (np.poly1d tables agree to 1e-5 at LGR p=3)."""
    col = mp.Collocation([3], "LGR")
    compD = col.get_composite_differentiation_matrix()
    compW = col.get_composite_quadrature_weights()
    taus = col.roots[3]
    assert col.tau0 == taus[0] and col.tau1 == taus[-1]
    Dn, wn = np.zeros((4, 4)), np.zeros(4)
    for j in range(4):
        pj = np.poly1d([1.0])
        for i in range(4):
            if i != j:
                pj *= np.poly1d([1, -taus[i]]) / (taus[j] - taus[i])
        Dn[:, j] = np.polyder(pj)(taus)
        wn[j] = np.polyint(pj)(1.0) - np.polyint(pj)(-1.0)
    assert abs(compD.toarray() - Dn).max() < 1e-5
    assert abs(compW.ravel() - wn).max() < 1e-5


def test_collocation_tables_match_oracle_at_arbitrary_points(mp):
    from oracle import collocation as oc

    taus = np.linspace(-0.9, 0.95, 7)
    for scheme, p in (("LGR", 7), ("LGL", 12), ("CGL", 20)):
        col = mp.Collocation([p], scheme)
        r = oc.roots(scheme, p)
        assert np.abs(col.get_interpolation_matrix(taus, p) - oc.interpolation_matrix(r, taus)).max() < 1e-11
        D1 = oc.diff_matrix(r, taus, 1)
        assert np.abs(col.get_diff_matrix(p, taus=taus) - D1).max() < 1e-10 * max(1, np.abs(D1).max())
        D2 = oc.diff_matrix(r, taus, 2)
        assert np.abs(col.get_diff_matrix(p, taus=taus, order=2) - D2).max() < 1e-9 * max(1, np.abs(D2).max())
        w = oc.quadrature_weights(r, -0.3, 0.8)
        assert np.abs(col.get_quadrature_weights(p, -0.3, 0.8) - w).max() < 1e-13
        ends = [np.array([-1.0, 1.0])]
        assert col.get_composite_interpolation_Dmatrix_at(ends, [p]).shape == (2, p + 1)


def test_structure_lengths_agree(mp):
    """tests/test_mpopt.py:333-407 on the reference's generic two-phase fixture."""
    from mpopt_b200.problems import generic_two_phase

    mpo = mp.mpopt(generic_two_phase())
    mpo.compute_numerical_approximation()
    N = mpo._Npoints
    assert len(mpo._taus) == len(set(mpo.poly_orders))
    assert mpo._compW.shape == (1, N) and mpo._compD.shape == (N, N)
    for p in mpo.poly_orders:
        assert len(mpo._taus[p]) == p + 1
    mpo.create_variables()
    for phase in range(2):
        Z, Zmin, Zmax = mpo.get_nlp_variables(phase)
        assert Z.shape[0] == Zmin.shape[0] == Zmax.shape[0] == mpo._optimization_vars_per_phase
        G, Gmin, Gmax, J = mpo.discretize_phase(phase)
        assert G.shape[0] == Gmin.shape[0] == Gmax.shape[0]
    E, Emin, Emax = mpo.get_event_constraints()
    assert len(E) == len(Emin) == len(Emax) == 3
    nlp_prob, nlp_bounds = mpo.create_nlp()
    assert nlp_prob["x"] == nlp_bounds["lbx"].shape[0] == nlp_bounds["ubx"].shape[0]
    assert nlp_bounds["lbg"].shape[0] == nlp_bounds["ubg"].shape[0] == mpo.transcription.n_g
    assert mpo.initialize_solution().shape[0] == mpo._optimization_vars_per_phase * 2


def test_bounds_and_initial_guess_match_oracle(mp):
    from mpopt_b200.problems import kitchen_sink, two_phase_schwartz
    from oracle.nlp import OracleNLP

    for ocp, K, p in ((kitchen_sink(), 4, [3, 2, 4, 3]), (two_phase_schwartz(), 3, 5)):
        mpo = mp.mpopt(ocp, K, p, "LGR")
        ora = OracleNLP(ocp, K, p, "LGR")
        for a, b in zip(mpo.transcription.bounds(), ora.bounds()):
            assert np.array_equal(a, b)
        assert np.allclose(mpo.initialize_solution(), ora.initialize_solution(), rtol=0, atol=1e-15)


def test_moon_lander_solve(mp):
    """tests/test_mpopt.py:416-428 / test_examples.py:47-48: the solution dict carries x and f; the optimum is the
    bang-bang cost 8.2477 (docs/source/notebooks/moon_lander.ipynb:185) up to discretisation."""
    from mpopt_b200.problems import moon_lander

    mpo, post = mp.solve(moon_lander(), n_segments=20, poly_orders=3, scheme="LGR", plot=False)
    sol = mpo.solve()
    for key in ("x", "f", "g", "lam_x", "lam_g", "lam_p"):
        assert key in sol
    for key in ("lbx", "lbg", "ubx", "ubg"):
        assert key in mpo.nlp_bounds
    assert abs(sol["f"] - 8.2477) < 5e-2
    zmin, zmax, gmin, gmax = mpo.transcription.bounds()
    assert (sol["g"] >= gmin - 1e-6).all() and (sol["g"] <= gmax + 1e-6).all()
    x, u, t, a = post.get_data()
    assert x.shape == (61, 2) and u.shape == (61, 1) and t.shape == (61, 1)
    assert abs(x[0, 0] - 10.0) < 1e-9 and abs(x[-1, 0]) < 1e-6 and abs(x[-1, 1]) < 1e-6


def test_analytic_solution(mp):
    """tests/test_mpopt.py:1090-1133: Chachuat ex. 3.10, x = -2t^2 + 6t + 1, u = 2(t - 1)."""
    from mpopt_b200.problems import chachuat_3_10

    mp.CollocationRoots._TAU_MIN = 0
    try:
        mpo = mp.mpopt(chachuat_3_10(), 1, 5)
        sol = mpo.solve(nlp_solver_options={"tol": 1e-14})
        post = mpo.process_results(sol, plot=False)
        x, u, t, _ = post.get_data()
        assert (abs(x - (-2 * t * t + 6 * t + 1)) < 1e-6).all()
        assert (abs(u - 2 * (t - 1)) < 1e-5).all()
    finally:
        mp.CollocationRoots._TAU_MIN = -1


def test_van_der_pol_solve(mp):
    """tests/test_mpopt.py:590-602 (CGL variant is the one that runs in the reference); optimum 2.8735
    (docs/source/notebooks/vanderpol.ipynb:191)."""
    from mpopt_b200.problems import van_der_pol

    mpo = mp.mpopt(van_der_pol(), 1, 15, "CGL")
    sol = mpo.solve()
    assert abs(sol["f"] - 2.8735) < 2e-2


def test_trust_constr_with_exact_hessian(mp):
    """The sparse solver path (trust-constr) fed by the Hessian kernel reaches the same optima as the reference's
    IPOPT runs: moon lander 8.2477 (moon_lander.ipynb:185); and it needs fewer iterations than with a BFGS model."""
    from mpopt_b200.problems import moon_lander

    mpo = mp.mpopt(moon_lander(), 8, 3, "LGR")
    sol = mpo.solve(nlp_solver_options={"method": "trust-constr", "max_iter": 300, "tol": 1e-8})
    it_exact = mpo.nlp_solver.stats["iter_count"]
    assert abs(sol["f"] - 8.2477) < 5e-2, (sol["f"], mpo.nlp_solver.stats)
    mpo2 = mp.mpopt(moon_lander(), 8, 3, "LGR")
    sol2 = mpo2.solve(nlp_solver_options={"method": "trust-constr", "max_iter": 300, "tol": 1e-8,
                                          "hessian_approximation": "limited-memory"})
    assert it_exact <= mpo2.nlp_solver.stats["iter_count"], (it_exact, mpo2.nlp_solver.stats)


def test_state_second_derivative(mp):
    """tests/test_mpopt.py:1136-1158: on the Chachuat solution x = -2 t^2 + 6 t + 1, u = 2 (t - 1) (tau in [0, 1], tf = 1)
    the second derivative of the state interpolant is -4 and that of the control 0."""
    from mpopt_b200.problems import chachuat_3_10

    mp.CollocationRoots._TAU_MIN = 0
    try:
        mpo = mp.mpopt(chachuat_3_10(), 1, 5)
        sol = mpo.solve(nlp_solver_options={"tol": 1e-14})
        taus = [mpo.collocation._taus_fn(deg)[1:-1] for deg in mpo.poly_orders]
        time, ddx, ddu = mpo.get_state_second_derivative_single_phase(sol, nodes=taus)
        assert all((abs(d + 4) < 1e-3).all() for d in ddx) and all((abs(d) < 1e-3).all() for d in ddu)
        ti, DDx, DDu = mpo.get_state_second_derivative(sol, nodes=[taus])
        assert len(DDx) == 1 and np.allclose(DDx[0][0], ddx[0])
    finally:
        mp.CollocationRoots._TAU_MIN = -1


def test_states_residuals_on_the_analytic_solution(mp):
    """tests/test_mpopt.py:730-760 style bound: on the Chachuat solution the state re-integrated from the dynamics agrees
    with the interpolated state (residual < 1e-3 everywhere), per phase / per segment lists like the reference's."""
    from mpopt_b200.problems import chachuat_3_10

    mp.CollocationRoots._TAU_MIN = 0
    try:
        mpo = mp.mpopt(chachuat_3_10(), 2, 4)
        sol = mpo.solve(nlp_solver_options={"tol": 1e-14})
        x_int, u_int, ti, res = mpo.get_states_residuals(sol)
        assert len(res) == 1 and len(res[0]) == 2
        for seg in res[0]:
            assert seg is not None and (np.abs(np.asarray(seg)) < 1e-3).all()
        t_all = np.concatenate(ti[0])
        x_all = np.vstack(x_int[0])[:, 0]
        assert (abs(x_all - (-2 * t_all ** 2 + 6 * t_all + 1)) < 1e-3).all()
    finally:
        mp.CollocationRoots._TAU_MIN = -1


def test_process_results_residual_flags_and_init_trajectories(mp):
    """mpopt.py:884-981 (residual_x / residual_dx attach the residual lists) and :857-882 (init_trajectories returns a
    callable (z, widths) -> (x, u, t, t0, tf, a) with scaled x, u)."""
    from mpopt_b200.problems import moon_lander

    mpo = mp.mpopt(moon_lander(), 5, 3, "LGR")
    sol = mpo.solve()
    post = mpo.process_results(sol, plot=False, residual_x=True, residual_dx=True)
    assert set(post.residuals) == {"t_x", "t_dx"}
    ti, res_dx = post.residuals["t_dx"]
    assert len(res_dx) == 1 and len(res_dx[0]) == 5
    ti2, res_x = post.residuals["t_x"]
    assert len(res_x[0]) == 5 and all(np.asarray(r).shape[1] == 2 for r in res_x[0] if r is not None)
    assert mpo.process_results(sol).residuals is None
    x, u, t, t0, tf, a = mpo.init_trajectories(0)(sol["x"], mpo._nlp_sw_params)
    assert x.shape == (16, 2) and u.shape == (16, 1) and t.shape == (16, 1)
    assert abs(t0[0]) < 1e-12 and abs(tf[0] - t[-1, 0]) < 1e-12 and abs(x[0, 0] - 10.0) < 1e-9
    # unequal widths move the interior time grid but not its ends
    w = np.array([0.1, 0.2, 0.3, 0.2, 0.2])
    x2, u2, t2, *_ = mpo.init_trajectories(0)(sol["x"], w)
    assert abs(t2[-1, 0] - t[-1, 0]) < 1e-12 and abs(t2[3, 0] - 0.1 * t[-1, 0]) < 1e-12
