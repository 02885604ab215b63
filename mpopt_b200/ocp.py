"""Optimal-control problem container with the reference's ``mp.OCP`` surface.

Mirrors the public attributes, defaults and helper methods of ``class OCP`` in
/root/reference/mpopt/mpopt.py:3378-3703 so that user scripts written for the
reference (dynamics / costs / constraints as Python callables, bounds and guesses
as numpy arrays indexed by phase) run unchanged.  Pure data: nothing here is on
the GPU path; the callables are traced once by :mod:`mpopt_b200.trace`.
"""
from __future__ import annotations

import numpy as np


def _per_phase(n_phases, width, value):
    return np.array([[value] * width for _ in range(n_phases)], dtype=float).reshape(n_phases, width)


class OCP:
    """Bolza-form multi-phase OCP (reference: mpopt.py:3378-3492).

    >>> ocp = OCP(n_states=2, n_controls=1)
    >>> ocp.dynamics[0] = lambda x, u, t: [x[1], u[0] - 1.5]
    >>> ocp.running_costs[0] = lambda x, u, t: u[0]
    >>> ocp.terminal_constraints[0] = lambda xf, tf, x0, t0: [xf[0], xf[1]]
    """

    # constraint-row bounds used by the transcription (mpopt.py:3392-3397)
    LB_DYNAMICS = 0
    UB_DYNAMICS = 0
    LB_PATH_CONSTRAINTS = -np.inf
    UB_PATH_CONSTRAINTS = 0
    LB_TERMINAL_CONSTRAINTS = 0
    UB_TERMINAL_CONSTRAINTS = 0

    def __init__(self, n_states=1, n_controls=1, n_phases=1, n_params=0, **kwargs):
        self.nx, self.nu, self.na, self.n_phases = n_states, n_controls, n_params, n_phases
        P, nx, nu, na = n_phases, n_states, n_controls, n_params

        # callables; `a` is optional so 3-/4-argument user lambdas both fit (mpopt.py:3426-3439)
        self.dynamics = [lambda x, u, t, a=None: [0] * self.nx] * P
        self.path_constraints = [lambda x, u, t, a=None: None] * P
        self.terminal_costs = [lambda xf, tf, x0, t0, a=None: 0] * P
        self.running_costs = [lambda x, u, t, a=None: 0] * P
        self.terminal_constraints = [lambda xf, tf, x0, t0, a=None: None] * P

        self.phase_links = [(i, i + 1) for i in range(P - 1)]  # :3442

        self.scale_x = np.ones(nx)
        self.scale_u = np.ones(nu)
        self.scale_a = np.ones(na)
        self.scale_t = 1.0

        # initial guess (:3451-3457)
        self.x00, self.xf0 = _per_phase(P, nx, 0.0), _per_phase(P, nx, 0.0)
        self.u00, self.uf0 = _per_phase(P, nu, 0.0), _per_phase(P, nu, 0.0)
        self.t00, self.tf0 = _per_phase(P, 1, 0.0), _per_phase(P, 1, 1.0)
        self.a0 = _per_phase(P, na, 0.0)

        # bounds (:3460-3472); phase 0 starts at t = 0
        self.lbx, self.ubx = _per_phase(P, nx, -np.inf), _per_phase(P, nx, np.inf)
        self.lbu, self.ubu = _per_phase(P, nu, -np.inf), _per_phase(P, nu, np.inf)
        self.lba, self.uba = _per_phase(P, na, -np.inf), _per_phase(P, na, np.inf)
        self.lbt0, self.ubt0 = _per_phase(P, 1, 0.0), _per_phase(P, 1, np.inf)
        self.ubt0[0] = 0.0
        self.lbtf, self.ubtf = _per_phase(P, 1, 0.0), _per_phase(P, 1, np.inf)

        # allowed state jump across phase links (:3475-3476)
        self.lbe, self.ube = _per_phase(P - 1, nx, 0.0), _per_phase(P - 1, nx, 0.0)

        # optional constraint blocks (:3479-3486)
        self.diff_u = np.array([0] * P)
        self.lbdu = np.array([-15 for _ in range(P)])
        self.ubdu = np.array([15 for _ in range(P)])
        self.midu = np.array([1] * P)
        self.du_continuity = np.array([0] * P)

        # post-processing defaults kept for API compatibility (:3489-3492)
        self.n_figures = 1
        self.phases_to_plot = [tuple(range(P))]
        self.plot_type = 1
        self.plot_interpolation_level = 3

    # ---- 4-argument adapters: user callables omit `a` when na == 0 (:3494-3571)
    def _adapt(self, fn):
        return (lambda *args: fn(*args[:-1])) if self.na == 0 else fn

    def get_dynamics(self, phase=0):
        return self._adapt(self.dynamics[phase])

    def get_path_constraints(self, phase=0):
        return self._adapt(self.path_constraints[phase])

    def get_running_costs(self, phase=0):
        return self._adapt(self.running_costs[phase])

    def get_terminal_constraints(self, phase=0):
        return self._adapt(self.terminal_constraints[phase])

    def get_terminal_costs(self, phase=0):
        return self._adapt(self.terminal_costs[phase])

    # ---- presence probes: evaluate on the numeric guess, "is not None" (:3573-3626)
    def has_path_constraints(self, phase=0):
        a = (self.a0[phase],) if self.na else ()
        return self.path_constraints[phase](self.x00[phase], self.u00[phase], self.t00[phase], *a) is not None

    def has_terminal_constraints(self, phase=0):
        a = (self.a0[phase],) if self.na else ()
        return (
            self.terminal_constraints[phase](self.xf0[phase], self.tf0[phase], self.x00[phase], self.t00[phase], *a)
            is not None
        )

    def validate(self):
        """Shape / ordering checks of mpopt.py:3628-3703."""
        P = self.n_phases
        assert P > 0
        for lst in (self.dynamics, self.running_costs, self.terminal_costs, self.path_constraints,
                    self.terminal_constraints):
            assert len(lst) == P
        for ph in range(P):
            x, u, t, a = self.x00[ph], self.u00[ph], self.t00[ph], self.a0[ph]
            assert len(self.get_dynamics(ph)(x, u, t, a)) == self.nx
            assert self.get_terminal_costs(ph)(x, t, x, t, a) is not None
            assert self.get_running_costs(ph)(x, u, t, a) is not None
            pc = self.get_path_constraints(ph)(x, u, t, a)
            tc = self.get_terminal_constraints(ph)(x, t, x, t, a)
            assert pc is None or len(pc) > 0
            assert tc is None or len(tc) > 0
        assert len(self.scale_x) == self.nx and len(self.scale_u) == self.nu and len(self.scale_a) == self.na
        for arr, w in ((self.x00, self.nx), (self.xf0, self.nx), (self.u00, self.nu), (self.uf0, self.nu),
                       (self.a0, self.na), (self.t00, 1), (self.tf0, 1), (self.lbx, self.nx), (self.ubx, self.nx),
                       (self.lbu, self.nu), (self.ubu, self.nu), (self.lba, self.na), (self.uba, self.na),
                       (self.lbt0, 1), (self.ubt0, 1), (self.lbtf, 1), (self.ubtf, 1)):
            assert np.shape(arr) == (P, w)
        assert self.lbe.shape[0] == P - 1 and self.ube.shape[0] == P - 1
        if P > 1:
            assert self.lbe.shape[1] == self.nx and self.ube.shape[1] == self.nx
        for ph in range(P):
            assert (self.lbx[ph] <= self.ubx[ph]).all() and (self.lbu[ph] <= self.ubu[ph]).all()
            assert (self.lba[ph] <= self.uba[ph]).all()
            assert self.lbt0[ph] <= self.ubt0[ph] and self.lbtf[ph] <= self.ubtf[ph]
            if ph < P - 1:
                assert (self.lbe[ph] <= self.ube[ph]).all()
