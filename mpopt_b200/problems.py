"""Optimal-control problems used by the tests, the benchmark and the AOT kernel registry.

Each function returns a fresh ``OCP``.  The definitions restate, as this package's own
fixtures, the problems the reference tests and BASELINE.json configurations are built on:
moon-lander (tests/test_mpopt.py:113-144), hyper-sensitive (:147-161), two-phase Schwartz
(:164-202), van-der-Pol (:205-227), Chachuat ex. 3.10 (:1090-1112), robot arm
(examples/singlephase/robot_arm.py:37-83) and the seeded synthetic 6-state/3-control
quadratic dynamics of SURVEY.md section 8(d).
"""
from __future__ import annotations

import numpy as np

from . import ca
from .ocp import OCP


def moon_lander():
    ocp = OCP(n_states=2, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [x[1], u[0] - 1.5]
    ocp.running_costs[0] = lambda x, u, t: u[0]
    ocp.terminal_constraints[0] = lambda xf, tf, x0, t0: [xf[0], xf[1]]
    ocp.tf0[0] = 4.0
    ocp.x00[0] = [10.0, -2.0]
    ocp.lbx[0] = [-20.0, -20.0]
    ocp.ubx[0] = [20.0, 20.0]
    ocp.lbu[0] = 0
    ocp.ubu[0] = 3
    ocp.lbtf[0], ocp.ubtf[0] = 3, 5
    ocp.validate()
    return ocp


def hyper_sensitive():
    ocp = OCP(n_states=1, n_controls=1, n_phases=1)
    ocp.dynamics[0] = lambda x, u, t: [-x[0] * x[0] * x[0] + u[0]]
    ocp.running_costs[0] = lambda x, u, t: 0.5 * (x[0] * x[0] + u[0] * u[0])
    ocp.terminal_constraints[0] = lambda xf, tf, x0, t0: [xf[0] - 1.0]
    ocp.x00[0] = 1
    ocp.lbtf[0] = ocp.ubtf[0] = 1000.0
    ocp.scale_t = 1 / 1000.0
    ocp.validate()
    return ocp


def two_phase_schwartz():
    ocp = OCP(n_states=2, n_controls=1, n_phases=2)

    def dynamics0(x, u, t):
        return [x[1], u[0] - 0.1 * (1.0 + 2.0 * x[0] * x[0]) * x[1]]

    ocp.dynamics = [dynamics0, dynamics0]
    ocp.path_constraints[0] = lambda x, u, t: [
        1.0 - 9.0 * (x[0] - 1) * (x[0] - 1) - (x[1] - 0.4) * (x[1] - 0.4) / (0.3 * 0.3)
    ]
    ocp.terminal_costs[1] = lambda xf, tf, x0, t0: 5 * (xf[0] * xf[0] + xf[1] * xf[1])
    ocp.x00[0] = [1, 1]
    ocp.x00[1] = [1, 1]
    ocp.xf0[0] = [1, 1]
    ocp.xf0[1] = [0, 0]
    ocp.lbx[0][1] = -0.8
    ocp.lbu[0], ocp.ubu[0] = -1, 1
    ocp.lbt0[0], ocp.ubt0[0] = 0, 0
    ocp.lbtf[0], ocp.ubtf[0] = 1, 1
    ocp.lbtf[1], ocp.ubtf[1] = 2.9, 2.9
    ocp.validate()
    return ocp


def van_der_pol():
    ocp = OCP(n_states=2, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [(1 - x[1] * x[1]) * x[0] - x[1] + u[0], x[0]]
    ocp.running_costs[0] = lambda x, u, t: x[0] * x[0] + x[1] * x[1] + u[0] * u[0]
    ocp.x00[0] = [0, 1]
    ocp.lbu[0] = -1.0
    ocp.ubu[0] = 1.0
    ocp.lbx[0][1] = -0.25
    ocp.lbtf[0] = 10.0
    ocp.ubtf[0] = 10.0
    ocp.validate()
    return ocp


def chachuat_3_10():
    """x' = 2(1-u), min int 0.5 u^2 - x; analytic solution x = -2t^2+6t+1, u = 2(t-1)."""
    ocp = OCP(n_states=1, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [2 * (1 - u[0])]
    ocp.running_costs[0] = lambda x, u, t: 0.5 * u[0] * u[0] - x[0]
    ocp.x00[0] = [1.0]
    ocp.lbtf[0] = 1.0
    ocp.ubtf[0] = 1.0
    ocp.validate()
    return ocp


def generic_two_phase():
    """The reference's structural test fixture (tests/test_mpopt.py:88-110): 2 states, 2 controls, 2 phases."""
    ocp = OCP(n_states=2, n_controls=2, n_phases=2)
    ocp.dynamics = [lambda x, u, t: [u[0], u[0]]] * 2
    ocp.path_constraints = [lambda x, u, t: [x[0] + 1, u[0]]] * 2
    ocp.running_costs = [lambda x, u, t: u[0]] * 2
    ocp.terminal_constraints = [lambda xf, tf, x0, t0: [-xf[0]]] * 2
    ocp.terminal_costs = [lambda xf, tf, x0, t0: tf] * 2
    for phase in range(2):
        ocp.lbu[phase], ocp.ubu[phase] = -1.0, 1.0
        ocp.lbtf[phase], ocp.ubtf[phase] = 1.0, 1.0
    ocp.validate()
    return ocp


def robot_arm():
    ocp = OCP(n_states=6, n_controls=3)

    def dynamics0(x, u, t):
        return [
            x[1],
            u[0] / 5.0,
            x[3],
            u[1] / (((5.0 - x[0]) ** 3 + x[0] ** 3) * ca.sin(x[4]) * ca.sin(x[4]) / 3.0),
            x[5],
            u[2] / (((5.0 - x[0]) ** 3 + x[0] ** 3) / 3.0),
        ]

    ocp.dynamics[0] = dynamics0
    ocp.terminal_costs[0] = lambda xf, tf, x0, t0: tf
    ocp.terminal_constraints[0] = lambda xf, tf, x0, t0: [
        xf[0] - 4.5, xf[1], xf[2] - 2.0 * np.pi / 3.0, xf[3], xf[4] - np.pi / 4.0, xf[5]]
    ocp.x00[0] = [4.5, 0, 0, 0, np.pi / 4.0, 0.0]
    ocp.xf0[0] = [4.5, 0, 2.0 * np.pi / 3.0, 0, np.pi / 4.0, 0.0]
    ocp.tf0[0] = 10
    ocp.lbu[0] = [-1.0, -1.0, -1.0]
    ocp.ubu[0] = [1.0, 1.0, 1.0]
    ocp.lbtf[0] = 10 - 3.0
    ocp.ubtf[0] = 10 + 3.0
    ocp.validate()
    return ocp


def synthetic_6_3():
    """Seeded dense quadratic dynamics, nx=6 nu=3: f_s = A_s.x + B_s.u + x_s (C_s.x), L = x.x + u.u."""
    rng = np.random.default_rng(6)
    A = rng.uniform(-1, 1, (6, 6))
    B = rng.uniform(-1, 1, (6, 3))
    C = rng.uniform(-1, 1, (6, 6))
    ocp = OCP(n_states=6, n_controls=3)

    def dynamics(x, u, t):
        return [
            sum(float(A[s, j]) * x[j] for j in range(6)) + sum(float(B[s, c]) * u[c] for c in range(3))
            + x[s] * sum(float(C[s, j]) * x[j] for j in range(6))
            for s in range(6)
        ]

    ocp.dynamics[0] = dynamics
    ocp.running_costs[0] = lambda x, u, t: sum(x[s] * x[s] for s in range(6)) + sum(u[c] * u[c] for c in range(3))
    ocp.lbu[0] = [-1.0] * 3
    ocp.ubu[0] = [1.0] * 3
    ocp.lbtf[0] = 1.0
    ocp.ubtf[0] = 1.0
    ocp.validate()
    return ocp


def kitchen_sink():
    """Everything at once: parameters, explicit time, scaling, path rows using t, Mayer term, slope rows.

    Not from the reference; exercises the branches the examples leave untouched."""
    ocp = OCP(n_states=3, n_controls=2, n_phases=2, n_params=2)

    def dyn(x, u, t, a):
        return [x[1] * a[0] + ca.sin(t) * u[0], -x[0] + u[1] * u[1] + a[1] * t, 0.5]

    ocp.dynamics = [dyn, lambda x, u, t, a: [u[0] - x[2] ** 3, ca.exp(-x[0] * x[0]) * a[0], x[1] / (1.0 + t * t)]]
    ocp.path_constraints[0] = lambda x, u, t, a: [x[0] * u[1] - t, ca.sqrt(1.0 + x[2] * x[2]) - a[1] - 3.0]
    ocp.running_costs = [lambda x, u, t, a: u[0] * u[0] + t * x[0] + a[0] * a[0],
                         lambda x, u, t, a: ca.cos(x[1]) * u[1] * u[1]]
    ocp.terminal_constraints[1] = lambda xf, tf, x0, t0, a: [xf[0] * xf[1] - a[0], tf - t0 - 2.0 + x0[2]]
    ocp.terminal_costs = [lambda xf, tf, x0, t0, a: tf * xf[2], lambda xf, tf, x0, t0, a: (xf[0] - x0[0]) ** 2 + a[1]]
    ocp.scale_x = np.array([2.0, 0.5, 4.0])
    ocp.scale_u = np.array([3.0, 0.25])
    ocp.scale_a = np.array([10.0, 0.1])
    ocp.scale_t = 0.5
    ocp.diff_u[:] = 1
    ocp.du_continuity[:] = 1
    for ph in range(2):
        ocp.lbu[ph], ocp.ubu[ph] = [-2.0, -np.inf], [2.0, np.inf]
        ocp.x00[ph] = [1.0, 0.5, -0.5]
        ocp.xf0[ph] = [0.5, 1.0, 0.5]
        ocp.a0[ph] = [0.3, 0.7]
        ocp.tf0[ph] = 2.0 + ph
        ocp.t00[ph] = 1.0 * ph
    ocp.validate()
    return ocp


#: problems whose node functors are compiled ahead of time into libmpx.so by build()
REGISTRY = {
    "moon_lander": moon_lander,
    "hyper_sensitive": hyper_sensitive,
    "two_phase_schwartz": two_phase_schwartz,
    "van_der_pol": van_der_pol,
    "chachuat_3_10": chachuat_3_10,
    "generic_two_phase": generic_two_phase,
    "robot_arm": robot_arm,
    "synthetic_6_3": synthetic_6_3,
    "kitchen_sink": kitchen_sink,
}

#: uniform polynomial degrees for which build() also compiles the degree-specialised g + jac_g kernel
#: (mpx_gjac4_kernel<PH, JAC, DEG>) of a registered problem; any other uniform degree is specialised at plan
#: creation through NVRTC when the plan is large, and runs the generic instance otherwise
AOT_DEGREES = {
    "synthetic_6_3": (15, 20),   # BASELINE.json headline (p=15) and the sharded configuration (p=20)
    "moon_lander": (15,),        # BASELINE.json configs[1]
    "two_phase_schwartz": (10,), # stand-in for BASELINE.json configs[4]
}
