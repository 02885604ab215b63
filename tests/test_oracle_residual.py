"""Oracle of the interpolation / residual path against the reference's own known answers
(/root/reference/tests/test_mpopt.py:663-675, :1161-1196) and against first principles."""
import numpy as np
import pytest

from oracle import residual as R
from oracle.nlp import OracleNLP


def test_interpolation_taus_known_answers():
    # tests/test_mpopt.py:663-675
    taus = R.interpolation_taus_on_original_grid(np.array([0, 0.5, 1]), [1])
    assert (abs(taus[0] - np.array([0.5, 1.0])) < 1e-6).all()
    taus = R.interpolation_taus_on_original_grid(np.array([0, 0.5, 1]), [0.5, 0.5])
    assert abs(taus[0][-1] - 1) < 1e-6 and abs(taus[1][-1] - 1) < 1e-6
    assert len(taus[0]) == 1 and len(taus[1]) == 1  # a boundary node belongs to the earlier segment, node 0 to nobody


def test_interpolated_time_grid_known_answers():
    # tests/test_mpopt.py:1161-1196
    t = np.array([0, 0.33, 1])
    assert (abs(R.interpolated_time_grid(t, [t], [2], 0, 1) - t) < 1e-6).all()
    t = np.array([0, 0.5, 1])
    assert (abs(R.interpolated_time_grid(t, [np.array([0, 1]), np.array([1])], [1, 1], 0, 1) - t) < 1e-6).all()
    assert (abs(R.interpolated_time_grid(t, [np.array([-1, 0, 1])], [2], -1, 1) - t) < 1e-6).all()


@pytest.mark.parametrize("grid", ["fixed", "mid-points", "spectral"])
def test_residual_grid_shapes(grid):
    # tests/test_mpopt.py:640-660: one array per segment, all inside [tau0, tau1]
    from mpopt_b200.problems import van_der_pol

    ora = OracleNLP(van_der_pol(), 3, [3, 5, 4], "LGR")
    taus = R.residual_grid_taus(ora, 0, grid)
    assert len(taus) == 3
    flat = np.concatenate(taus)
    assert flat.min() >= ora.tau0 - 1e-12 and flat.max() <= ora.tau1 + 1e-12  # (t - c) / w rounds a hair past 1
    assert R.residual_grid_taus(ora, 0, "do-not-know-any") is None


def test_residual_vanishes_on_an_exact_polynomial_solution():
    """x' = u with u = 2 t, x = t^2 on [0, 2]: a degree-2 trajectory is represented exactly, so the interpolated state
    is t^2 everywhere and the dynamics residual D_I X - h f is zero at any point (and equals the defect rows at nodes)."""
    from mpopt_b200 import OCP

    ocp = OCP(n_states=1, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [u[0]]
    ocp.validate()
    ora = OracleNLP(ocp, 4, [3, 2, 4, 3], "LGL")
    p = np.array([0.1, 0.4, 0.3, 0.2])
    z = np.zeros(ora.n_z)
    z[ora.colT0(0)], z[ora.colTF(0)] = 0.0, 2.0
    _, t, _, _ = ora._time_grid(0, 0.0, 2.0, p)
    z[: ora.N] = t ** 2
    z[ora.N: 2 * ora.N] = 2 * t
    taus = [np.array([-0.7, 0.1, 0.9]), np.array([]), np.array([0.0]), np.array([-1.0, 1.0])]
    Xi, Ui, ti, DXi, DUi = R.interpolate_phase(ora, z, p, 0, taus)
    assert np.allclose(Xi[:, 0], ti ** 2, atol=1e-12) and np.allclose(Ui[:, 0], 2 * ti, atol=1e-12)
    ti2, res, F, n = R.dynamics_residuals_phase(ora, z, p, 0, taus)
    assert n == [3, 0, 1, 2] and np.allclose(res, 0.0, atol=1e-11)
    tis, ress = R.dynamics_residuals(ora, z, p, nodes=[taus])
    assert ress[0][1] is None and ress[0][0].shape == (3, 1)


def test_second_derivative_of_an_exact_polynomial_solution():
    """mpopt.py:1285-1358 on x = t^2, u = 2 t over [0, 2] with unequal segments: d2x/dtau2 = 2 (dt/dtau)^2 = 2 (h_k / 1)^2
    with dt/dtau = h_k = (tf - t0)/delta * w_k, and d2u/dtau2 = 0 -- the known answer the reference checks in the same way
    for x = -2 t^2 + 6 t + 1 (tests/test_mpopt.py:1136-1158)."""
    from mpopt_b200 import OCP

    ocp = OCP(n_states=1, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [u[0]]
    ocp.validate()
    ora = OracleNLP(ocp, 3, [3, 2, 4], "LGR")
    p = np.array([0.2, 0.5, 0.3])
    z = np.zeros(ora.n_z)
    z[ora.colT0(0)], z[ora.colTF(0)] = 0.0, 2.0
    _, t, _, _ = ora._time_grid(0, 0.0, 2.0, p)
    z[: ora.N], z[ora.N: 2 * ora.N] = t ** 2, 2 * t
    taus = [np.array([-0.5, 0.25]), np.array([0.0]), np.array([-1.0, 0.3, 1.0])]
    ti, ddx, ddu = R.second_derivatives_phase(ora, z, p, 0, taus)
    h = (2.0 - 0.0) / 2.0 * np.repeat(p, [2, 1, 3])
    assert np.allclose(ddx[:, 0], 2 * h ** 2, atol=1e-10) and np.allclose(ddu, 0.0, atol=1e-10)
    assert ti.shape == (6,)


def test_state_residual_by_quadrature_known_answers():
    """mpopt.py:989-1076.  (i) On an exact polynomial solution (x' = u, x = t^2) re-integrating the dynamics from the
    segment start reproduces the interpolated state: residual 0.  (ii) With the control perturbed by a constant c the
    re-integrated state drifts by c * (t - t_segment_start): the residual is minus that, a hand-checkable answer."""
    from mpopt_b200 import OCP

    ocp = OCP(n_states=1, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [u[0]]
    ocp.validate()
    ora = OracleNLP(ocp, 3, [3, 2, 4], "LGR")
    p = np.array([0.2, 0.5, 0.3])
    z = np.zeros(ora.n_z)
    z[ora.colT0(0)], z[ora.colTF(0)] = 0.0, 2.0
    _, t, _, _ = ora._time_grid(0, 0.0, 2.0, p)
    z[: ora.N], z[ora.N: 2 * ora.N] = t ** 2, 2 * t
    taus = [np.array([-0.5, 0.25, 0.8]), np.array([]), np.array([-0.9, 0.0, 0.3, 1.0])]
    xint, res, ti = R.states_from_dynamics_phase(ora, z, p, 0, taus)
    assert np.allclose(xint[:, 0], ti ** 2, atol=1e-12) and np.allclose(res, 0.0, atol=1e-12)
    c = 0.37
    z2 = z.copy()
    z2[ora.N: 2 * ora.N] += c
    xint2, res2, ti2 = R.states_from_dynamics_phase(ora, z2, p, 0, taus)
    t_start = np.repeat(t[ora.seg_start], [3, 0, 4])
    assert np.allclose(res2[:, 0], -c * (ti2 - t_start), atol=1e-12)
