"""``mp.mpopt`` / ``mp.solve``: the reference's transcription + solve driver surface over the GPU evaluators.

Keeps the constructor, the public attributes (``n_segments, poly_orders, colloc_scheme, _ocp, _Npoints,
nlp_bounds, nlp_solver, _nlp_sw_params``) and the methods of the reference's ``class mpopt``
(/root/reference/mpopt/mpopt.py:31-1573) that belong to the hot path: ``compute_numerical_approximation``,
``create_nlp``, ``create_solver``, ``solve``, ``initialize_solution``, ``get_segment_width_parameters``,
``get_solver_warm_start_input_parameters``, ``get_nlp_variables``, ``discretize_phase``,
``get_event_constraints``.  What CasADi built symbolically there is a ``Transcription`` here: the user's Python
callables are traced once and the four NLP evaluators run as CUDA kernels.

Out of scope: the adaptive subclasses, residual post-processing and plotting.
``process_results`` returns a light object with ``get_data`` / ``get_trajectories`` only.
"""
from __future__ import annotations

import copy
import time

import numpy as np

from .collocation import Collocation, CollocationRoots
from .nlp import Transcription
from .solver import ScipyNlpSolver


class mpopt:
    _GRID_TYPE = "fixed"
    _MAX_GRID_POINTS = 15
    _MUTE_ = False

    def __init__(self, problem, n_segments: int = 1, poly_orders=[9], scheme: str = "LGR", **kwargs):
        self.n_segments = n_segments
        self.poly_orders = [poly_orders] * n_segments if isinstance(poly_orders, (int, np.integer)) else poly_orders
        self._ocp = copy.deepcopy(problem)  # later edits of the caller's OCP are ignored, as in the reference (:77)
        self.colloc_scheme = scheme
        self.device = int(kwargs.get("device", 0))
        self.reset_mpopt()

    def reset_mpopt(self):
        assert len(self.poly_orders) == self.n_segments  # :83
        self._Npoints = sum(self.poly_orders) + 1
        self._collocation_approximation_computed = False
        self._variables_created = False
        self._nlpsolver_initialized = False
        self._tr = None
        self.grid_type = [self._GRID_TYPE for _ in range(self._ocp.n_phases)]
        self.max_grid_points = [self._MAX_GRID_POINTS for _ in range(self._ocp.n_phases)]

    # ------------------------------------------------------------------ tables
    def compute_numerical_approximation(self, scheme: str = None):
        scheme = self.colloc_scheme if scheme is None else scheme
        self.collocation = Collocation(self.poly_orders, scheme, device=self.device)
        self._compD = self.collocation.get_composite_differentiation_matrix()
        self._compW = self.collocation.get_composite_quadrature_weights()
        self._taus = self.collocation.roots
        self.tau0, self.tau1 = self.collocation.tau0, self.collocation.tau1
        self._collocation_approximation_computed = True

    # ------------------------------------------------------------------ transcription
    @property
    def transcription(self) -> Transcription:
        if self._tr is None:
            self._tr = Transcription(self._ocp, self.n_segments, self.poly_orders, self.colloc_scheme,
                                     tau_min=float(CollocationRoots._TAU_MIN), tau_max=float(CollocationRoots._TAU_MAX),
                                     device=self.device)
            o = self._ocp
            self._optimization_vars_per_phase = self._Npoints * (o.nx + o.nu) + o.na + 2
            self._variables_created = True
        return self._tr

    def create_variables(self):
        self.transcription  # noqa: B018  (layout is fixed when the plan is created)

    def get_nlp_variables(self, phase: int):
        """(Z, Zmin, Zmax) of one phase; Z is the index range of the phase in the decision vector."""
        tr = self.transcription
        zmin, zmax, _, _ = tr.bounds()
        n = tr.layout.nvar
        sl = slice(phase * n, (phase + 1) * n)
        return np.arange(sl.start, sl.stop), zmin[sl], zmax[sl]

    def _phase_rows(self, phase):
        L = self.transcription.layout
        start = L.phases[phase].gF
        stop = L.phases[phase + 1].gF if phase + 1 < len(L.phases) else L.g_events
        return start, stop

    def discretize_phase(self, phase: int):
        """(G, Gmin, Gmax, J): row indices of the phase's constraints, their bounds and the objective callable."""
        tr = self.transcription
        _, _, gmin, gmax = tr.bounds()
        a, b = self._phase_rows(phase)
        return np.arange(a, b), gmin[a:b], gmax[a:b], tr.f

    def get_event_constraints(self):
        tr = self.transcription
        if self._ocp.n_phases < 2:
            return ([], [], [])
        _, _, gmin, gmax = tr.bounds()
        o, n = self._ocp, len(self._ocp.phase_links)
        sizes = [n * o.nx, n * o.nu, n]
        a = tr.layout.g_events
        E, Emin, Emax = [], [], []
        for sz in sizes:
            E.append(np.arange(a, a + sz)), Emin.append(gmin[a:a + sz]), Emax.append(gmax[a:a + sz])
            a += sz
        return (E, Emin, Emax)

    def create_nlp(self):
        """(nlp_problem, nlp_bounds): the evaluators and the bound vectors (reference: symbolic f/x/g/p, :574-639)."""
        if not self._collocation_approximation_computed:
            self.compute_numerical_approximation()
        tr = self.transcription
        self.Zmin, self.Zmax, self.Gmin, self.Gmax = tr.bounds()
        nlp_prob = {"f": tr.f, "grad_f": tr.grad_f, "g": tr.g, "jac_g": tr.jac_g, "x": tr.n_z, "p": tr.n_p,
                    "transcription": tr}
        nlp_bounds = {"lbg": self.Gmin, "ubg": self.Gmax, "lbx": self.Zmin, "ubx": self.Zmax}
        return (nlp_prob, nlp_bounds)

    def initialize_solution(self):
        return self.transcription.initial_guess()

    def init_solution_per_phase(self, phase: int):
        n = self.transcription.layout.nvar
        return self.initialize_solution()[phase * n:(phase + 1) * n]

    def get_segment_width_parameters(self, solution=None):
        return [1.0 / self.n_segments] * (self.n_segments * self._ocp.n_phases)  # :723

    # ------------------------------------------------------------------ solver
    def create_solver(self, solver: str = "ipopt", options={}):
        nlp_problem, self.nlp_bounds = self.create_nlp()
        opts = {"ipopt.max_iter": 2000, "ipopt.acceptable_tol": 1e-4, "ipopt.print_level": 0, "ipopt.sb": "yes",
                "print_time": 0} if solver == "ipopt" else {}  # the reference's defaults (:743-749)
        opts.update(options)
        self.nlp_solver = ScipyNlpSolver(nlp_problem["transcription"], opts)
        self._nlpsolver_initialized = True

    def get_solver_warm_start_input_parameters(self, solution=None):
        pairs = {"x": "x0", "x0": "x0", "lam_x": "lam_x0", "lam_x0": "lam_x0", "lam_g": "lam_g0", "lam_g0": "lam_g0"}
        inputs = {}
        if solution is not None:
            for k in solution:
                if k in pairs:
                    inputs[pairs[k]] = solution[k]
        if "x0" not in inputs:
            inputs["x0"] = self.initialize_solution()
        return inputs

    def solve(self, initial_solution=None, reinitialize_nlp=False, solver="ipopt", nlp_solver_options={},
              mpopt_options={}, **kwargs):
        if not self._MUTE_:
            print("\n *********** MPOPT Summary ********** \n")
        t0 = time.monotonic()
        if (not self._nlpsolver_initialized) or reinitialize_nlp:
            self.create_solver(solver=solver, options=nlp_solver_options)
        self._nlp_sw_params = mpopt_options["nlp_sw_params"] if "nlp_sw_params" in mpopt_options else \
            self.get_segment_width_parameters(initial_solution)
        inputs = self.get_solver_warm_start_input_parameters(initial_solution)
        inputs["p"] = self._nlp_sw_params
        t1 = time.monotonic()
        solution = self.nlp_solver(**inputs, **self.nlp_bounds)
        t2 = time.monotonic()
        if not self._MUTE_:
            print(" Optimal cost (J): ", solution["f"], "\n")
            print(f" Solved in {round((t2 - t0) * 1e3, 3)} ms")
            print(f" \t OCP transcription time  : {round((t1 - t0) * 1e3, 3)} ms")
            print(f" \t NLP solution time       : {round((t2 - t1) * 1e3, 3)} ms")
        return solution

    # ------------------------------------------------------------------ results
    def process_results(self, solution, plot: bool = False, scaling: bool = False, residual_x=False, residual_dx=False):
        return post_process(solution, self, scaling)

    def validate(self):
        pass


class post_process:
    """Trajectory extraction only (the reference's plotting / residual machinery is out of scope)."""

    def __init__(self, solution, mpo, scaling=False):
        self.solution, self.mpo, self.scaling = solution, mpo, scaling
        self.phases = list(range(mpo._ocp.n_phases))

    def get_trajectories(self, phase: int = 0):
        """(x, u, t, a) of one phase, unscaled unless ``scaling`` (mpopt.py:1639-1667)."""
        mpo = self.mpo
        tr, o = mpo.transcription, mpo._ocp
        L, N = tr.layout, tr.N
        z = np.asarray(self.solution["x"], dtype=float).reshape(-1)
        off = phase * L.nvar
        X = z[off:off + o.nx * N].reshape(o.nx, N).T
        U = z[off + o.nx * N:off + (o.nx + o.nu) * N].reshape(o.nu, N).T
        T0, TF = z[L.colT0(phase)] / o.scale_t, z[L.colTF(phase)] / o.scale_t
        A = z[L.colT0(phase) + 2:off + L.nvar]
        w = np.asarray(getattr(mpo, "_nlp_sw_params", mpo.get_segment_width_parameters()), float)[phase * tr.K:(phase + 1) * tr.K]
        delta = tr.tau1 - tr.tau0
        t = np.empty(N)
        t[0], acc = T0, T0
        for k, p in enumerate(tr.poly_orders):
            r = tr.tables(p)[0]
            h = (TF - T0) / delta * w[k]
            s0 = int(L.seg_start[k])
            t[s0 + 1:s0 + p + 1] = acc + h * (r[1:] - tr.tau0)
            acc = acc + h * delta
        t = t.reshape(-1, 1)
        if self.scaling:
            return X, U, t, A
        return X / o.scale_x, U / o.scale_u, t, A / o.scale_a

    def get_original_data(self, phases=[]):
        phases = phases or self.phases
        parts = [self.get_trajectories(ph) for ph in phases]
        return tuple(np.vstack([p[i] for p in parts]) if i < 3 else np.concatenate([np.atleast_1d(p[i]) for p in parts])
                     for i in range(4))

    def get_data(self, phases=[], interpolate: bool = False):
        return self.get_original_data(phases)


def solve(ocp, n_segments=1, poly_orders=9, scheme="LGR", plot=False, solve_dict=dict(), residual_x=False,
          residual_dx=False):
    """One-liner of the reference (mpopt.py:4279-4308): returns (optimizer, post-processor)."""
    mpo = mpopt(ocp, n_segments=n_segments, poly_orders=poly_orders, scheme=scheme)
    solution = mpo.solve(**solve_dict)
    post = mpo.process_results(solution, plot=False)
    return (mpo, post)
