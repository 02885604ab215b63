"""The numpy oracle against the numbers the reference itself holds: the optimal objectives stored in its executed
notebooks, at the notebooks' own discretisations (tests/anchors.py).  This is what pins the oracle's VALUES -- nodes,
D, weights, layout, functions, first and second derivatives -- to the reference beyond the p = 1 known answers."""
import numpy as np
import pytest

import anchors as A


@pytest.mark.parametrize("problem,K,p,scheme,ref,where,rtol,seen", A.ANCHORS, ids=[f"{a[0]}-{a[3]}" for a in A.ANCHORS])
def test_oracle_optimum_matches_reference_notebook(problem, K, p, scheme, ref, where, rtol, seen):
    A.check_anchor("oracle", problem, K, p, scheme, ref, rtol)


@pytest.mark.parametrize("problem,K,p,scheme,atol,where", A.ANCHORS_ZERO, ids=[a[3] for a in A.ANCHORS_ZERO])
def test_oracle_two_phase_schwartz_optimum_is_zero(problem, K, p, scheme, atol, where):
    from mpopt_b200.problems import REGISTRY

    r = A.Evaluators("oracle", REGISTRY[problem](), K, p, scheme).solve(tol=1e-10)
    assert r.success and abs(r.f) <= atol


def test_oracle_delta3_mayer_optimum_matches_reference_notebook():
    A.check_delta3("oracle")


def test_q2_band():
    """Quirk Q2 made quantitative: with quadrature weights produced by an ODE solver at IDAS's default tolerances
    instead of the exact ones, the moon-lander optimum moves by a few 1e-6 relative -- the size of the gap between
    this package's optimum and the reference's stored one -- and the stored value lies within that band."""
    from scipy.integrate import solve_ivp

    import oracle.collocation as oc
    from mpopt_b200.problems import REGISTRY

    problem, K, p, scheme, ref = "moon_lander", 10, 6, "LGR", 8.2477255075783038
    exact = A.Evaluators("oracle", REGISTRY[problem](), K, p, scheme).solve(tol=1e-10).f
    orig = oc.quadrature_weights

    def ode_weights(r, a, b):
        def ell(j, t):
            v = 1.0
            for i in range(len(r)):
                if i != j:
                    v *= (t - r[i]) / (r[j] - r[i])
            return v
        return np.array([solve_ivp(lambda t, x, j=j: [ell(j, t)], (a, b), [0.0], method="BDF", rtol=1e-6, atol=1e-8).y[0, -1]
                         for j in range(len(r))])

    oc.quadrature_weights = ode_weights
    try:
        shifted = A.Evaluators("oracle", REGISTRY[problem](), K, p, scheme).solve(tol=1e-10).f
    finally:
        oc.quadrature_weights = orig
    shift, gap = abs(shifted - exact) / ref, abs(exact - ref) / ref
    assert 1e-7 < shift < 2e-5, shift           # the weights' integration error is visible at this level ...
    assert gap <= 2.0 * shift, (gap, shift)     # ... and accounts for the gap to the stored objective
