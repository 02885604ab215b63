"""The CUDA path (through the C ABI) against the numbers the reference itself holds: optimal objectives stored in its
executed notebooks, at the notebooks' own (n_segments, poly_orders, scheme) -- tests/anchors.py.  Every evaluation of
these solves (f, grad_f, g, jac_g, hess_l) is a kernel launch; nothing of oracle/ is involved."""
import pytest

import anchors as A

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("problem,K,p,scheme,ref,where,rtol,seen", A.ANCHORS, ids=[f"{a[0]}-{a[3]}" for a in A.ANCHORS])
def test_cuda_optimum_matches_reference_notebook(libmpx, problem, K, p, scheme, ref, where, rtol, seen):
    r = A.check_anchor("cuda", problem, K, p, scheme, ref, rtol)
    assert r.iter > 3


@pytest.mark.parametrize("problem,K,p,scheme,atol,where", A.ANCHORS_ZERO, ids=[a[3] for a in A.ANCHORS_ZERO])
def test_cuda_two_phase_schwartz_optimum_is_zero(libmpx, problem, K, p, scheme, atol, where):
    from mpopt_b200.problems import REGISTRY

    r = A.Evaluators("cuda", REGISTRY[problem](), K, p, scheme).solve(tol=1e-10)
    assert r.success and abs(r.f) <= atol


def test_cuda_delta3_mayer_optimum_matches_reference_notebook(libmpx):
    A.check_delta3("cuda")


def test_public_api_reproduces_the_notebook_runs(libmpx):
    """mp.solve(ocp, n_segments, poly_orders, scheme) -- the call the notebooks make -- returns their objectives."""
    import mpopt_b200.mp as mp
    from mpopt_b200.problems import moon_lander, van_der_pol

    mpo, post = mp.solve(moon_lander(), n_segments=10, poly_orders=6, scheme="LGR", plot=False,
                         solve_dict={"nlp_solver_options": {"tol": 1e-10}})
    x, u, t, _ = post.get_data()
    assert x.shape == (61, 2) and abs(x[-1, 0]) < 1e-8 and abs(x[-1, 1]) < 1e-8  # soft landing
    mpo2 = mp.mpopt(van_der_pol(), 1, 25, "LGL")
    sol2 = mpo2.solve(nlp_solver_options={"tol": 1e-10})
    assert abs(sol2["f"] - 2.8734849959084205) <= 3e-6 * 2.8734849959084205
    mpo3 = mp.mpopt(moon_lander(), 2, 30, "LGL")
    sol3 = mpo3.solve(nlp_solver_options={"tol": 1e-10})
    assert abs(sol3["f"] - 8.2425586640613506) <= 5e-7 * 8.2425586640613506
    assert mpo3.nlp_solver.stats["method"] == "ipm" and mpo3.nlp_solver.stats["success"]
