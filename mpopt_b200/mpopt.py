"""``mp.mpopt`` / ``mp.solve``: the reference's transcription + solve driver surface over the GPU evaluators.

Keeps the constructor, the public attributes (``n_segments, poly_orders, colloc_scheme, _ocp, _Npoints,
nlp_bounds, nlp_solver, _nlp_sw_params``) and the methods of the reference's ``class mpopt``
(/root/reference/mpopt/mpopt.py:31-1573) that belong to the hot path: ``compute_numerical_approximation``,
``create_nlp``, ``create_solver``, ``solve``, ``initialize_solution``, ``get_segment_width_parameters``,
``get_solver_warm_start_input_parameters``, ``get_nlp_variables``, ``discretize_phase``,
``get_event_constraints``.  What CasADi built symbolically there is a ``Transcription`` here: the user's Python
callables are traced once and the four NLP evaluators run as CUDA kernels.

Also here, from the "next" rows of SURVEY.md section 8f: the interpolation / dynamics-residual path
(``get_residual_grid_taus``, ``interpolate_single_phase``, ``get_dynamics_residuals*``, evaluated on the GPU).
Out of scope: the adaptive subclasses and plotting; ``process_results`` returns a light object with ``get_data`` /
``get_trajectories`` only.
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

    def _current_widths(self):
        """Width fractions of the last solve (``_nlp_sw_params``), equal widths before the first one."""
        sw = getattr(self, "_nlp_sw_params", None)
        if sw is None or len(sw) == 0:
            sw = [1.0 / self.n_segments] * (self.n_segments * self._ocp.n_phases)
        return np.asarray(sw, dtype=float)

    def _solution_widths(self, z):
        """Width fractions the time grid of solution vector ``z`` is built with: here the parameters of the last solve
        (mpopt.py:873-896 evaluates the trajectories with ``_nlp_sw_params``); ``mpopt_adaptive`` reads them out of ``z``."""
        return self._current_widths()

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

    # ------------------------------------------------------------------ interpolation / residuals (mpopt.py:1152-1573)
    @staticmethod
    def compute_interpolation_taus_corresponding_to_original_grid(nodes_req, seg_widths, tau0=0, tau1=1):
        """Global target nodes -> per-segment local taus (mpopt.py:1205-1237): a node on a segment boundary belongs to
        the earlier segment, the first node to nobody."""
        nodes_req = np.asarray(nodes_req, dtype=float)
        csw = np.append(0, np.cumsum(seg_widths))
        assert abs(csw[-1] - 1) < 1e-6
        scaled = 0 + (1 - 0) / (tau1 - tau0) * (nodes_req - tau0)
        taus = [None] * len(seg_widths)
        for i, seg in enumerate(seg_widths):
            t = scaled[scaled > csw[i]]
            t = t[t <= csw[i + 1]]
            taus[i] = tau0 + (tau1 - tau0) / (1 - 0) * ((t - csw[i]) / seg - 0)
        return taus

    @staticmethod
    def get_interpolated_time_grid(t_orig, taus, poly_orders, tau0, tau1):
        """mpopt.py:1544-1573 (returns a column, like the reference's DM)."""
        t_orig = np.asarray(t_orig, dtype=float).reshape(-1)
        t_seg = [t_orig[0]] + [t_orig[sum(poly_orders[: i + 1])] for i in range(len(poly_orders))]
        grid = [t_seg[i] + (t_seg[i + 1] - t_seg[i]) * (0 + (1 - 0) / (tau1 - tau0) * (np.asarray(taus[i], float) - tau0))
                for i in range(len(t_seg) - 1)]
        return np.concatenate(grid).reshape(-1, 1)

    def get_residual_grid_taus(self, phase: int = 0, grid_type: str = None):
        """Non-collocation points per segment: "fixed", "mid-points", "spectral" (mpopt.py:1152-1203)."""
        if not self._collocation_approximation_computed:
            self.compute_numerical_approximation()
        if grid_type is None:
            grid_type = self.grid_type[phase]
        sw = self._current_widths()
        if grid_type == "fixed":
            n_nodes = max(sum(self.poly_orders) + 2, self._MAX_GRID_POINTS + 2)
            target = np.linspace(self.tau0, self.tau1, n_nodes)
            taus = self.compute_interpolation_taus_corresponding_to_original_grid(
                target, sw[self.n_segments * phase: self.n_segments * (phase + 1)], tau0=self.tau0, tau1=self.tau1)
            taus[0] = taus[0][:-1]
            return taus
        if grid_type == "mid-points":
            return [(np.asarray(self.collocation._taus_fn(d))[:-1] + np.asarray(self.collocation._taus_fn(d))[1:]) / 2.0
                    for d in self.poly_orders]
        if grid_type == "spectral":
            return [np.array(self.collocation._taus_fn(self._MAX_GRID_POINTS + 2)[1:-1]) for _ in self.poly_orders]
        return None

    def interpolate_single_phase(self, solution, phase: int = 0, target_nodes=None, grid_type=None, options={}):
        """(Xi, Ui, ti, a, DXi, DUi, target_nodes, t0, tf) at per-segment taus, evaluated on the GPU (mpopt.py:1489-1542)."""
        if target_nodes is None:
            target_nodes = self.get_residual_grid_taus(phase=phase, grid_type=grid_type)
        tr, o = self.transcription, self._ocp
        z = np.asarray(solution["x"], dtype=float).reshape(-1)
        sw = self._current_widths()
        r = tr.residuals(z, sw, phase, target_nodes)
        L = tr.layout
        a = z[L.colT0(phase) + 2: L.colT0(phase) + 2 + o.na]
        t0, tf = z[L.colT0(phase)] / o.scale_t, z[L.colTF(phase)] / o.scale_t
        self._last_residuals = r
        return (r["xi"], r["ui"], r["ti"].reshape(-1, 1), a, r["dxi"], r["dui"], target_nodes, np.atleast_1d(t0),
                np.atleast_1d(tf))

    def get_dynamics_residuals_single_phase(self, solution, phase: int = 0, target_nodes=None):
        """(ti, residual, h Sx f) per segment, None / [] for a segment without points (mpopt.py:1428-1487)."""
        xi, ui, ti, a, dxi, dui, taus, t0, tf = self.interpolate_single_phase(solution, phase, target_nodes)
        res = self._last_residuals["res"]
        n = [len(t) for t in taus]
        off = np.concatenate([[0], np.cumsum(n)]).astype(int)
        K = self.n_segments
        res_seg = [res[off[k]: off[k + 1]] if n[k] else None for k in range(K)]
        dyn_seg = [(dxi - res)[off[k]: off[k + 1]] if n[k] else None for k in range(K)]
        ti_seg = [ti[off[k]: off[k + 1]] if n[k] else [] for k in range(K)]
        return ti_seg, res_seg, dyn_seg

    def get_dynamics_residuals(self, solution, nodes=None, grid_type=None, residual_type=None, plot=False, fig=None,
                               axs=None):
        """Residual of the dynamics at non-collocation points, per phase and segment (mpopt.py:1360-1426)."""
        residuals, ti = [None] * self._ocp.n_phases, [None] * self._ocp.n_phases
        for phase in range(self._ocp.n_phases):
            target = nodes[phase] if nodes is not None else self.get_residual_grid_taus(
                phase, grid_type=self.grid_type[phase] if grid_type is None else grid_type)
            ti[phase], residuals[phase], dyn = self.get_dynamics_residuals_single_phase(solution, phase, target)
            if residual_type == "relative":
                mx = np.zeros(self._ocp.nx)
                for d in dyn:
                    if d is not None:
                        mx = np.maximum(mx, np.abs(d).max(axis=0))
                residuals[phase] = [r / mx if r is not None else None for r in residuals[phase]]
        return ti, residuals

    def compute_states_from_solution_dynamics(self, solution, phase: int = 0, nodes=None):
        """(xint, u, ti, residual) per segment: the states re-integrated from the dynamics at the target points by
        quadrature, and their difference to the interpolated states, evaluated on the GPU (mpopt.py:989-1076)."""
        target = nodes if nodes is not None else self.get_residual_grid_taus(phase=phase, grid_type=self.grid_type[phase])
        z = np.asarray(solution["x"], dtype=float).reshape(-1)
        r = self.transcription.state_residuals(z, self._current_widths(), phase, target)
        off = np.concatenate([[0], np.cumsum(r["counts"])]).astype(int)
        K = self.n_segments
        xint, uu, ti, res = [None] * K, [None] * K, [None] * K, [None] * K
        for k in range(K):
            if off[k] == off[k + 1]:
                continue
            sl = slice(off[k], off[k + 1])
            xint[k], uu[k], ti[k], res[k] = r["xint"][sl], r["ui"][sl], r["ti"][sl], list(r["res_x"][sl])
        return xint, uu, ti, res

    def get_states_residuals(self, solution, phases=None, nodes=None, residual_type=None, plot=False, fig=None, axs=None):
        """mpopt.py:1078-1150: the above for the given phases; ``residual_type="relative"`` divides by the largest
        |xint| per state."""
        P = self._ocp.n_phases
        x_int, u_int, residuals, ti = [None] * P, [None] * P, [None] * P, [None] * P
        for phase in (range(P) if phases is None else phases):
            target = nodes[phase] if nodes is not None else self.get_residual_grid_taus(phase, grid_type=self.grid_type[phase])
            x_int[phase], u_int[phase], ti[phase], residuals[phase] = self.compute_states_from_solution_dynamics(
                solution, phase, nodes=target)
            if residual_type == "relative":
                mx = np.zeros(self._ocp.nx)
                for seg in x_int[phase]:
                    if seg is not None:
                        mx = np.maximum(mx, np.abs(np.asarray(seg)).max(axis=0))
                residuals[phase] = [np.asarray(r_) / mx if r_ is not None else None for r_ in residuals[phase]]
        return x_int, u_int, ti, residuals

    def get_state_second_derivative_single_phase(self, solution, phase: int = 0, nodes=None, grid_type: str = None,
                                                 residual_type: str = None):
        """(ti, ddx, ddu) per segment: second tau-derivative of the state / control interpolants at the given local
        taus, evaluated on the GPU (mpopt.py:1285-1358); None for a segment without points."""
        target = nodes if nodes is not None else self.get_residual_grid_taus(phase=phase, grid_type=self.grid_type[phase])
        z = np.asarray(solution["x"], dtype=float).reshape(-1)
        r = self.transcription.second_derivatives(z, self._current_widths(), phase, target)
        off = np.concatenate([[0], np.cumsum(r["counts"])]).astype(int)
        K = self.n_segments
        ti, ddx, ddu = [None] * K, [None] * K, [None] * K
        for k in range(K):
            if off[k] == off[k + 1]:
                continue
            ddx[k], ddu[k] = r["ddxi"][off[k]: off[k + 1]], r["ddui"][off[k]: off[k + 1]]
            if residual_type == "relative":
                ddx[k], ddu[k] = ddx[k] / ddx[k].max(), ddu[k] / ddu[k].max()
            ti[k] = r["ti"][off[k]: off[k + 1]]
        return ti, ddx, ddu

    def get_state_second_derivative(self, solution, grid_type="spectral", nodes=None, plot=False, fig=None, axs=None):
        """mpopt.py:1238-1283: the above for every phase."""
        P = self._ocp.n_phases
        ti, DDx, DDu = [None] * P, [None] * P, [None] * P
        for phase in range(P):
            target = nodes[phase] if nodes is not None else self.get_residual_grid_taus(phase, grid_type=grid_type)
            ti[phase], DDx[phase], DDu[phase] = self.get_state_second_derivative_single_phase(solution, phase, nodes=target)
        return ti, DDx, DDu

    # ------------------------------------------------------------------ results
    def init_trajectories(self, phase: int = 0):
        """Callable ``(z, seg_widths) -> (x, u, t, t0, tf, a)`` of one phase: scaled x / u, times in the OCP's units
        (the reference returns a CasADi Function with this signature, mpopt.py:857-882)."""
        def trajectories(z, seg_widths=None):
            post = post_process({"x": z}, self, scaling=True)
            widths = None if seg_widths is None or len(seg_widths) == 0 else np.asarray(seg_widths, dtype=float)
            x, u, t, a = post.get_trajectories(phase, widths=widths)
            return x, u, t, np.atleast_1d(t[0, 0]), np.atleast_1d(t[-1, 0]), a
        return trajectories

    def process_results(self, solution, plot: bool = False, scaling: bool = False, residual_x: bool = False,
                        residual_dx: bool = False):
        """Post-processor of a solution (mpopt.py:884-981).  ``residual_x`` / ``residual_dx`` evaluate the state and
        dynamics residuals on the GPU and attach them as ``post.residuals = {"t_x": [ti, res_x], "t_dx": [tdx, res_dx]}``
        (the reference's ``options["residuals"]``); plotting is out of scope, so ``plot`` is ignored."""
        post = post_process(solution, self, scaling)
        post.residuals = None
        if residual_x or residual_dx:
            post.residuals = {}
            if residual_x:
                _, _, ti, res_x = self.get_states_residuals(solution)
                post.residuals["t_x"] = [ti, res_x]
            if residual_dx:
                tdx, res_dx = self.get_dynamics_residuals(solution)
                post.residuals["t_dx"] = [tdx, res_dx]
        return post

    def validate(self):
        pass


class post_process:
    """Trajectory extraction only (the reference's plotting / residual machinery is out of scope)."""

    def __init__(self, solution, mpo, scaling=False):
        self.solution, self.mpo, self.scaling = solution, mpo, scaling
        self.phases = list(range(mpo._ocp.n_phases))

    def __getattr__(self, name):
        # plot_phases / plot_x / plot_u / plot_residuals ... (mpopt.py:1838-2270) need matplotlib: out of scope, say so
        if name.startswith("plot"):
            raise NotImplementedError(f"post_process.{name}: plotting (matplotlib) is outside the accelerated path; "
                                      "use get_data() / get_data(interpolate=True) and plot the arrays")
        raise AttributeError(name)

    def get_trajectories(self, phase: int = 0, widths=None):
        """(x, u, t, a) of one phase, unscaled unless ``scaling`` (mpopt.py:1639-1667); ``widths``: segment-width
        fractions of all phases to build the time grid with (default: those of the last solve)."""
        mpo = self.mpo
        tr, o = mpo.transcription, mpo._ocp
        L, N = tr.layout, tr.N
        z = np.asarray(self.solution["x"], dtype=float).reshape(-1)
        off = phase * L.nvar
        X = z[off:off + o.nx * N].reshape(o.nx, N).T
        U = z[off + o.nx * N:off + (o.nx + o.nu) * N].reshape(o.nu, N).T
        T0, TF = z[L.colT0(phase)] / o.scale_t, z[L.colTF(phase)] / o.scale_t
        A = z[L.colT0(phase) + 2:L.colT0(phase) + 2 + o.na]
        w = (mpo._solution_widths(z) if widths is None else np.asarray(widths, float))[phase * tr.K:(phase + 1) * tr.K]
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

    _INTERPOLATION_NODES_PER_SEG = 50

    def get_interpolation_taus(self, n: int = 75, taus_orig=None, method: str = "uniform"):
        """mpopt.py:1690-1709."""
        tr = self.mpo.transcription
        if method == "uniform" or taus_orig is None:
            return np.linspace(tr.tau0, tr.tau1, n)
        return self.get_non_uniform_interpolation_grid(taus_orig, n)

    @staticmethod
    def get_non_uniform_interpolation_grid(taus_orig, n: int = 75):
        """Insert mid-points until the grid has n points, at most six passes (mpopt.py:1711-1738)."""
        taus = np.asarray(taus_orig, dtype=float)
        for _ in range(6):
            if len(taus) >= n:
                break
            fine = np.empty(2 * len(taus) - 1)
            fine[0::2], fine[1::2] = taus, 0.5 * (taus[:-1] + taus[1:])
            taus = fine
        return taus

    def get_interpolated_data(self, phases, taus=[]):
        """(x, u, t, a) on a finer grid: Lagrange interpolation inside every segment, evaluated by the residual
        kernel on the GPU (mpopt.py:1767-1826; the reference multiplies with a dense composite matrix)."""
        mpo = self.mpo
        tr, o = mpo.transcription, mpo._ocp
        if not len(taus):
            taus = [self.get_interpolation_taus(n=self._INTERPOLATION_NODES_PER_SEG)[1:] for _ in tr.poly_orders]
            taus[0] = np.append(tr.tau0, taus[0])
        z = np.asarray(self.solution["x"], dtype=float).reshape(-1)
        sw = mpo._current_widths()
        xs, us, ts, As = [], [], [], []
        for phase in phases:
            r = tr.residuals(z, sw, phase, taus, derivatives=False)
            t_orig = self.get_trajectories(phase)[2]
            ts.append(np.asarray(mpopt.get_interpolated_time_grid(t_orig, taus, tr.poly_orders, tr.tau0, tr.tau1)).reshape(-1))
            xs.append(r["xi"] if self.scaling else r["xi"] / o.scale_x)
            us.append(r["ui"] if self.scaling else r["ui"] / o.scale_u)
            As.append(np.atleast_1d(self.get_trajectories(phase)[3]))
        return np.vstack(xs), np.vstack(us), np.hstack(ts), np.hstack(As)

    def get_data(self, phases=[], interpolate: bool = False):
        phases = phases or self.phases
        return self.get_interpolated_data(phases) if interpolate else self.get_original_data(phases)


def solve(ocp, n_segments=1, poly_orders=9, scheme="LGR", plot=False, solve_dict=dict(), residual_x=False,
          residual_dx=False):
    """One-liner of the reference (mpopt.py:4279-4308): returns (optimizer, post-processor)."""
    mpo = mpopt(ocp, n_segments=n_segments, poly_orders=poly_orders, scheme=scheme)
    solution = mpo.solve(**solve_dict)
    post = mpo.process_results(solution, plot=False, residual_x=residual_x, residual_dx=residual_dx)
    return (mpo, post)
