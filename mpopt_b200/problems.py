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


def delta3_launch_vehicle(drag=0.0):
    """Four-phase ascent of the Delta III to a geostationary transfer orbit: 7 states (position, velocity, mass in an
    Earth-centred inertial frame), 3 controls (thrust direction), path rows |u| = 1 and r >= Re, orbital-element
    terminal constraints, mass jumps at the stage separations, heavy scaling.

    Restates the OCP of the reference's examples/Multi-phase/multistage_launch_vehicle.py:35-296 (the problem behind
    docs/source/notebooks/multi_stage_launch_vehicle_ascent.ipynb, whose IPOPT banner -- 474 variables, 374 equalities,
    276 inequalities at one segment of degree 11 per phase, :466-471 -- pins the layout of a 4-phase NLP with events).
    ``drag`` is the example's ``param``: 0 switches the aerodynamic term off, and the 0 * D products then fold out of
    the Jacobian pattern as they do in CasADi."""
    Re, omega, mu = 6378145.0, 7.29211585e-5, 3.986012e14
    rho0, scale_h, area_cd = 1.225, 7200.0, 4 * np.pi * 0.5
    lat = 28.5 * np.pi / 180.0
    r_pad = np.array([Re * np.cos(lat), 0.0, Re * np.sin(lat)])
    v_pad = omega * np.array([-r_pad[1], r_pad[0], 0.0])
    m_lift, m_payload = 301454.0, 4164.0
    srb_prop, srb_dry = 17010.0, 19290.0 - 17010.0
    first_prop, first_dry = 95550.0, 104380.0 - 95550.0
    second_prop, second_dry = 16820.0, 19300.0 - 16820.0
    srb_burn, first_burn, second_burn = 75.2, 261.0, 700.0
    thrust = [6 * 628500.0 + 1083100.0, 3 * 628500.0 + 1083100.0, 1083100.0, 110094.0]
    flow = [6 * srb_prop / srb_burn + first_prop / first_burn, 3 * srb_prop / srb_burn + first_prop / first_burn,
            first_prop / first_burn, second_prop / second_burn]

    def norm3(v):
        return ca.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])

    def make_dynamics(T, mdot):
        def f(x, u, t):
            r, v, m = x[:3], x[3:6], x[6]
            rm = norm3(r)
            vrel = ca.vertcat(v[0] + r[1] * omega, v[1] - r[0] * omega, v[2])
            rho = rho0 * ca.exp(-(rm - Re) / scale_h)
            D = -rho / (2 * m) * area_cd * norm3(vrel) * vrel
            grav = -mu / (rm * rm * rm) * r
            return [x[3], x[4], x[5]] + [T / m * u[i] + drag * D[i] + grav[i] for i in range(3)] + [-mdot]
        return f

    def path(x, u, t):
        uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
        return [uu - 1, -uu + 1, -norm3(x) / Re + 1]

    a_req, e_req, i_req = 24361140.0, 0.7308, 28.5 * np.pi / 180.0
    node_req, argp_req = 269.8 * np.pi / 180.0, 130.5 * np.pi / 180.0

    def orbit(x, t, x0, t0):
        """Orbital elements of the final state minus the targets."""
        h = ca.vertcat(x[1] * x[5] - x[4] * x[2], x[3] * x[2] - x[0] * x[5], x[0] * x[4] - x[1] * x[3])
        n = ca.vertcat(-h[1], h[0], 0)
        r = norm3(x)
        e = ca.vertcat(1 / mu * (x[4] * h[2] - x[5] * h[1]) - x[0] / r,
                       1 / mu * (x[5] * h[0] - x[3] * h[2]) - x[1] / r,
                       1 / mu * (x[3] * h[1] - x[4] * h[0]) - x[2] / r)
        e_mag = norm3(e)
        v_mag = ca.sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5])
        a = -mu / (v_mag * v_mag - 2.0 * mu / r)
        inc = ca.acos(h[2] / ca.sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]))
        n_mag = ca.sqrt(n[0] * n[0] + n[1] * n[1])
        node = 2 * np.pi - ca.acos(n[0] / n_mag)
        argp = ca.acos((n[0] * e[0] + n[1] * e[1]) / (n_mag * e_mag))
        return [(a - a_req) / Re, e_mag - e_req, inc - i_req, node - node_req, argp - argp_req]

    ocp = OCP(n_states=7, n_controls=3, n_phases=4)
    ocp.dynamics = [make_dynamics(T, q) for T, q in zip(thrust, flow)]
    ocp.path_constraints = [path] * 4
    ocp.terminal_costs[3] = lambda xf, tf, x0, t0: -xf[-1] / m_lift
    ocp.terminal_constraints[3] = orbit
    vs = np.sqrt(mu / Re)
    ocp.scale_x = [1 / Re] * 3 + [1 / vs] * 3 + [1 / m_lift]
    ocp.scale_t = vs / Re

    # guesses: straight line from the pad to the perigee of the target orbit, masses from the burn schedule
    p_orb = a_req * (1.0 - e_req * e_req)
    rp = p_orb / (1.0 + e_req)
    cn, sn, cp, sp, ci, si = (np.cos(node_req), np.sin(node_req), np.cos(argp_req), np.sin(argp_req), np.cos(i_req),
                              np.sin(i_req))
    rot = np.array([[cn * cp - sn * sp * ci, -cn * sp - sn * cp * ci, sn * si],
                    [sn * cp + cn * sp * ci, -sn * sp + cn * cp * ci, -cn * si],
                    [sp * si, cp * si, ci]])
    r_end = rot @ np.array([rp, 0.0, 0.0])
    v_end = rot @ (np.sqrt(mu / p_orb) * np.array([0.0, e_req + 1.0, 0.0]))
    tk = [0.0, srb_burn, 2 * srb_burn, first_burn, 924.0]
    start = np.concatenate([r_pad, v_pad, [m_lift]])
    end = np.concatenate([r_end, v_end, [m_payload + second_dry]])
    at = lambda t: start + (end - start) / (tk[4] - tk[0]) * (t - tk[0])
    xs = [start.copy(), at(tk[1]), at(tk[2]), at(tk[3])]
    xe = [at(tk[1]), at(tk[2]), at(tk[3]), end.copy()]
    xe[0][-1] = xs[0][-1] - (6 * srb_prop + first_prop / tk[3] * tk[1])
    xs[1][-1] = xe[0][-1] - 6 * srb_dry
    xe[1][-1] = xs[1][-1] - (3 * srb_prop + first_prop / tk[3] * (tk[2] - tk[1]))
    xs[2][-1] = xe[1][-1] - 3 * srb_dry
    xe[2][-1] = xs[2][-1] - first_prop / tk[3] * (tk[3] - tk[2])
    xs[3][-1] = xe[2][-1] - first_dry
    ocp.x00, ocp.xf0 = np.array(xs), np.array(xe)
    ocp.u00 = np.array([[1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 1, 0]], dtype=float)
    ocp.uf0 = np.array([[0, 1, 0]] * 4, dtype=float)
    ocp.t00 = np.array([[t] for t in tk[:4]])
    ocp.tf0 = np.array([[t] for t in tk[1:]])
    box_lo, box_hi = [-2 * Re] * 3 + [-10000.0] * 3, [2 * Re] * 3 + [10000.0] * 3
    ocp.lbx = np.array([box_lo + [xe[k][-1]] for k in range(4)])
    ocp.ubx = np.array([box_hi + [xs[k][-1]] for k in range(4)])
    ocp.lbu, ocp.ubu = np.array([[-1.0] * 3] * 4), np.array([[1.0] * 3] * 4)
    ocp.lbt0 = ocp.ubt0 = np.array([[t] for t in tk[:4]])
    ocp.lbtf = np.array([[tk[1]], [tk[2]], [tk[3]], [tk[4] - 100]])
    ocp.ubtf = np.array([[tk[1]], [tk[2]], [tk[3]], [tk[4] + 100]])
    jumps = [-6 * srb_dry, -3 * srb_dry, -first_dry]
    ocp.lbe = np.array([[0.0] * 6 + [j] for j in jumps])
    ocp.ube = ocp.lbe.copy()
    ocp.validate()
    return ocp


def falcon9_launcher(drag=0.0, dyn_pressure=0.0, glide_slope=0.0):
    """Falcon 9 to orbit with booster return: 7 states, 4 controls (thrust direction + throttle), 3 phases linked as a
    FORK -- ascent (0) -> second stage to orbit (1) and ascent (0) -> booster return (2) -- with different path rows
    per phase (3, 3, 5), two of phase 2's rows constant while the dynamic-pressure and glide-slope switches are off.

    Restates the OCP of the reference's examples/Multi-phase/falcon9_launcher.py:35-304; the IPOPT banner of
    docs/source/notebooks/falcon9_to_orbit.ipynb:480-485 (956 variables, 746 equalities, 641 inequalities at five
    segments of degree 6) pins the fork's event rows and the mid-point rows whose bounds coincide (fixed throttle)."""
    Re, mu, g0 = 6378145.0, 3.986012e14, 9.80665
    rho0, scale_h, area_cd = 1.225, 7200.0, 4 * np.pi * 0.5
    lat = 28.5 * np.pi / 180.0
    pad = np.array([Re * np.cos(lat), 0.0, Re * np.sin(lat)])
    m_lift = 431.6e3 + 107.5e3
    m_final = 107.5e3 - 103.5e3
    booster_dry = 431.6e3 - 409.5e3
    start = np.concatenate([pad, 7.29211585e-5 * np.array([0.1, 0.1, 0.1]), [m_lift]])
    q_max, isp = 80e3, 340.0
    thrust = [9 * 934.0e3, 934.0e3, 934.0e3]

    def norm3(v):
        return ca.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])

    def make_dynamics(T):
        def f(x, u, t):
            r, v, m = x[:3], x[3:6], x[6]
            rm = norm3(r)
            rho = rho0 * ca.exp(-(rm - Re) / scale_h)
            D = -rho / (2 * m) * area_cd * norm3(v) * v
            grav = -mu / (rm * rm * rm) * r
            return ([x[3], x[4], x[5]] + [T * u[3] / m * u[i] + drag * D[i] + grav[i] for i in range(3)]
                    + [-T * u[3] / (isp * g0)])
        return f

    def path_ascent(x, u, t):
        uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
        return [uu - 1, -uu + 1, -norm3(x) / Re + 1]

    def path_return(x, u, t):
        rho = rho0 * ca.exp(-(norm3(x) - Re) / scale_h)
        v_sq = x[3] * x[3] + x[4] * x[4] + x[5] * x[5]
        rel = ca.vertcat(x[0] - start[0], x[1] - start[1], x[2] - start[2])
        uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
        cone = norm3(rel) * np.cos(80.0 * np.pi / 180.0) - (rel[0] * start[0] + rel[1] * start[1] + rel[2] * start[2]) / np.sqrt(
            start[0] ** 2 + start[1] ** 2 + start[2] ** 2)
        return [dyn_pressure * 0.5 * rho * v_sq / q_max - 1.0, uu - 1, -uu + 1, -norm3(x) / Re + 1, glide_slope * cone]

    a_req, e_req, i_req = 6593145.0, 0.0076, 28.5 * np.pi / 180.0
    node_req, argp_req = 269.8 * np.pi / 180.0, 130.5 * np.pi / 180.0

    def orbit(x, t, x0, t0):
        h = ca.vertcat(x[1] * x[5] - x[4] * x[2], x[3] * x[2] - x[0] * x[5], x[0] * x[4] - x[1] * x[3])
        n = ca.vertcat(-h[1], h[0], 0)
        r = norm3(x)
        e = ca.vertcat(1 / mu * (x[4] * h[2] - x[5] * h[1]) - x[0] / r,
                       1 / mu * (x[5] * h[0] - x[3] * h[2]) - x[1] / r,
                       1 / mu * (x[3] * h[1] - x[4] * h[0]) - x[2] / r)
        e_mag = norm3(e)
        v_mag = ca.sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5])
        a = -mu / (v_mag * v_mag - 2.0 * mu / r)
        inc = ca.acos(h[2] / ca.sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]))
        n_mag = ca.sqrt(n[0] * n[0] + n[1] * n[1])
        node = 2 * np.pi - ca.acos(n[0] / n_mag)
        argp = ca.acos((n[0] * e[0] + n[1] * e[1]) / (n_mag * e_mag))
        return [(a - a_req) / Re, e_mag - e_req, inc - i_req, node - node_req, argp - argp_req]

    vs = np.sqrt(mu / Re)

    def back_to_pad(x, t, x_0, t_0):
        return [(x[i] - start[i]) / Re for i in range(3)] + [(x[i] - start[i]) / vs for i in range(3, 6)]

    ocp = OCP(n_states=7, n_controls=4, n_phases=3)
    ocp.dynamics = [make_dynamics(T) for T in thrust]
    ocp.path_constraints = [path_ascent, path_ascent, path_return]
    ocp.terminal_costs[1] = lambda xf, tf, x0, t0: -xf[6] / m_lift
    ocp.terminal_constraints[1] = orbit
    ocp.terminal_constraints[2] = back_to_pad
    ocp.scale_x = np.array([1 / Re] * 3 + [1 / vs] * 3 + [1 / m_lift])
    ocp.scale_t = vs / Re

    p_orb = a_req * (1.0 - e_req * e_req)
    cn, sn, cp, sp, ci, si = (np.cos(node_req), np.sin(node_req), np.cos(argp_req), np.sin(argp_req), np.cos(i_req),
                              np.sin(i_req))
    rot = np.array([[cn * cp - sn * sp * ci, -cn * sp - sn * cp * ci, sn * si],
                    [sn * cp + cn * sp * ci, -sn * sp + cn * cp * ci, -cn * si],
                    [sp * si, cp * si, ci]])
    end = np.concatenate([rot @ np.array([p_orb / (1.0 + e_req), 0.0, 0.0]),
                          rot @ (np.sqrt(mu / p_orb) * np.array([0.0, e_req + 1.0, 0.0])), [m_final]])
    t_sep, t_orbit, t_land = 131.4, 453.4, 569.7
    at_sep = start + (end - start) / t_orbit * t_sep
    burnt = 9 * 934e3 / (isp * g0) * t_sep
    sep_end = at_sep.copy()
    sep_end[-1] = start[-1] - burnt
    left_in_booster = 409.5e3 - burnt
    second_start = at_sep.copy()
    second_start[-1] = sep_end[-1] - (booster_dry + left_in_booster)
    ocp.x00 = np.array([start, second_start, sep_end])
    ocp.xf0 = np.array([sep_end, end, start])
    ocp.u00 = np.array([[1, 0, 0, 1.0], [1, 0, 0, 1], [0, 1, 0, 1]])
    ocp.uf0 = np.array([[0, 1, 0, 1.0], [0, 1, 0, 1], [1, 0, 0, 0.5]])
    ocp.t00 = np.array([[0.0], [t_sep], [t_sep]])
    ocp.tf0 = np.array([[t_sep], [t_orbit], [t_land]])
    box_lo, box_hi = [-2 * Re] * 3 + [-10000.0] * 3, [2 * Re] * 3 + [10000.0] * 3
    ocp.lbx = np.array([box_lo + [sep_end[-1]], box_lo + [end[-1]], box_lo + [booster_dry]])
    ocp.ubx = np.array([box_hi + [start[-1]], box_hi + [107.5e3], box_hi + [sep_end[-1] - 107.5e3]])
    ocp.lbu = np.array([[-1.0, -1.0, -1.0, 1.0], [-1.0, -1.0, -1.0, 1.0], [-1.0, -1.0, -1.0, 0.38]])
    ocp.ubu = np.array([[1.0] * 4] * 3)
    ocp.lbt0 = ocp.ubt0 = np.array([[0.0], [t_sep], [t_sep]])
    ocp.lbtf = np.array([[t_sep], [t_orbit - 50], [t_land - 100]])
    ocp.ubtf = np.array([[t_sep], [t_orbit + 50], [t_land + 100]])
    ocp.lbe = np.array([[0.0] * 6 + [-(booster_dry + left_in_booster)], [0.0] * 6 + [-107.5e3]])
    ocp.ube = ocp.lbe.copy()
    ocp.phase_links = [(0, 1), (0, 2)]
    ocp.validate()
    return ocp


def alp_rider():
    """Betts' alp rider (stiff 4-state system, a path constraint that depends on time through four Gaussian peaks);
    restates examples/singlephase/Betts/alpr01_alp_rider.py:34-90, which tests/test_examples.py:38-50 solves."""
    ocp = OCP(n_states=4, n_controls=2)
    ocp.dynamics[0] = lambda x, u, t: [-10 * x[0] + u[0] + u[1], -2 * x[1] + u[0] + 2 * u[1],
                                       -3 * x[2] + 5 * x[3] + u[0] - u[1], 5 * x[2] - 3 * x[3] + u[0] + 3 * u[1]]
    ocp.terminal_constraints[0] = lambda xf, tf, x0, t0: [xf[0] - 2.0, xf[1] - 3.0, xf[2] - 1.0, xf[3] + 2]
    ocp.running_costs[0] = lambda x, u, t: (100 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3])
                                            + 0.01 * (u[0] * u[0] + u[1] * u[1]))
    peaks = ((3.0, 12, 3), (3.0, 10, 6), (3.0, 6, 10), (8.0, 4, 15))
    ocp.path_constraints[0] = lambda x, u, t: [
        sum(amp * ca.exp(-sharp * (t - at) * (t - at)) for amp, sharp, at in peaks) + 0.01
        - x[0] * x[0] - x[1] * x[1] - x[2] * x[2] - x[3] * x[3]]
    ocp.x00[0] = [2.0, 1.0, 2.0, 1.0]
    ocp.xf0[0] = [2.0, 3.0, 1.0, -2.0]
    ocp.tf0[0] = 20
    ocp.lbtf[0] = ocp.ubtf[0] = 20.0
    ocp.validate()
    return ocp


def mine_opt():
    """Optimal ore extraction (the Wikipedia optimal-control example): cost u^2 / x - p u divides by the state;
    restates examples/singlephase/mine_opt_wiki.py:30-52."""
    ocp = OCP(n_states=1, n_controls=1)
    ocp.dynamics[0] = lambda x, u, t: [-u[0]]
    ocp.running_costs[0] = lambda x, u, t: u[0] * u[0] / x[0] - 1 * u[0]
    ocp.x00[0] = [1.0]
    ocp.lbx[0] = 0
    ocp.ubx[0] = 1.0
    ocp.lbtf[0] = ocp.ubtf[0] = 1.0
    ocp.validate()
    return ocp


def dae_van_der_pol():
    """Van der Pol with the state bound turned into a path constraint on a free parameter (n_params = 1, callables
    with the 4-argument signature); restates examples/singlephase/dae_vdp.py:28-62."""
    ocp = OCP(n_states=2, n_controls=1, n_params=1)
    ocp.dynamics[0] = lambda x, u, t, a: [(1 - x[1] * x[1]) * x[0] - x[1] + u[0], x[0]]
    ocp.running_costs[0] = lambda x, u, t, a: x[0] * x[0] + x[1] * x[1] + u[0] * u[0]
    ocp.path_constraints[0] = lambda x, u, t, a: [a[0] - x[1]]
    ocp.x00[0] = [0, 1]
    ocp.lbu[0], ocp.ubu[0] = -1.0, 1.0
    ocp.lba[0], ocp.uba[0] = 0.25, 0.5
    ocp.lbx[0][1] = -0.25
    ocp.lbtf[0] = ocp.ubtf[0] = 10.0
    ocp.validate()
    return ocp


#: the reference's remaining example problems (tests/test_examples.py:38-50): not compiled ahead of time, so they
#: exercise the run-time (NVRTC) route
EXAMPLES = {"alp_rider": alp_rider, "mine_opt": mine_opt, "dae_van_der_pol": dae_van_der_pol}


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
    "delta3_launch_vehicle": delta3_launch_vehicle,
    "falcon9_launcher": falcon9_launcher,
}

#: uniform polynomial degrees for which build() also compiles the degree-specialised g + jac_g kernel
#: (mpx_gjac4_kernel<PH, JAC, DEG>) of a registered problem; any other uniform degree is specialised at plan
#: creation through NVRTC when the plan is large, and runs the generic instance otherwise
AOT_DEGREES = {
    "synthetic_6_3": (15, 20),   # BASELINE.json headline (p=15) and the sharded configuration (p=20)
    "moon_lander": (15,),        # BASELINE.json configs[1]
    "two_phase_schwartz": (10,), # stand-in for BASELINE.json configs[4]
}
