"""``mp.mpopt_h_adaptive``: iterative segment-width refinement around the GPU evaluators.

The caller of the hot path in the reference (/root/reference/mpopt/mpopt.py:2273-2874): the NLP is
transcribed ONCE, the segment widths are its parameter vector ``p`` (mpopt.py:152, :631), and every refinement pass
re-solves the same device plan with new widths and re-evaluates the dynamics residual between the collocation nodes
(``mpx_eval_residuals``, the N3 kernel).  Nothing is re-traced or re-compiled between passes.

Kept from the reference: class-level knobs (``_TOL_RESIDUAL``, ``_TOL_SEG_WIDTH_CHANGE``, ``_THRESHOLD_SLOPE`` ...),
``solve(max_iter=, mpopt_options={"method", "sub_method"})``, ``iter_count`` / ``iter_info``, the three width
updates -- ``residual/merge_split`` (:2663-2707), ``residual/equal_area`` (:2637-2660) with its 0.4 / 0.6 blend
(:2588-2591) and ``control_slope`` (:2709-2874) -- and the stopping rules of the outer loop (:2393-2460).
The plotting hooks (``plot_residual_evolution``) are accepted and ignored: matplotlib is out of scope.
"""
from __future__ import annotations

import itertools
import time

import numpy as np

from .mpopt import mpopt


def _seg_peak(res):
    """max |residual| of one segment (0 for a segment without evaluation points, mpopt.py:2567-2572)."""
    return float(np.abs(np.asarray(res)).max()) if res is not None and np.size(res) else 0.0


class mpopt_h_adaptive(mpopt):
    _SEG_WIDTH_MIN = 1e-5
    _SEG_WIDTH_MAX = 1
    _TOL_SEG_WIDTH_CHANGE = 0.05
    _TOL_RESIDUAL = 1e-2
    _DEFAULT_METHOD = "residual"
    _DEFAULT_SUB_METHOD = "equal_area"
    _THRESHOLD_SLOPE = 1e-1

    def __init__(self, problem, n_segments: int = 1, poly_orders=[9], scheme: str = "LGR", **kwargs):
        super().__init__(problem, n_segments=n_segments, poly_orders=poly_orders, scheme=scheme, **kwargs)
        P = self._ocp.n_phases
        self.lbh = [self._SEG_WIDTH_MIN] * P
        self.ubh = [self._SEG_WIDTH_MAX] * P
        self.tol_residual = [self._TOL_RESIDUAL] * P
        self.fig, self.axs = None, None
        self.plot_residual_evolution = False

    def _say(self, *a):
        if not self._MUTE_:
            print(*a)

    # ------------------------------------------------------------------ outer loop (mpopt.py:2331-2472)
    def solve(self, initial_solution=None, reinitialize_nlp=False, solver="ipopt", nlp_solver_options={},
              mpopt_options={}, max_iter: int = 10, **kwargs):
        self._say("\n *********** MPOPT H-Adaptive Summary ********** \n")
        tic = time.monotonic()
        if (not self._nlpsolver_initialized) or reinitialize_nlp:
            opts = dict(nlp_solver_options)
            opts.setdefault("ipopt.print_level", 0)
            self.create_solver(solver=solver, options=opts)
        options = dict(mpopt_options) if mpopt_options else {"method": self._DEFAULT_METHOD,
                                                             "sub_method": self._DEFAULT_SUB_METHOD}
        tol = min(self.tol_residual)
        self.iter_count, self.iter_info = 0, {}
        widths_prev = []
        widths_next, max_error = self.get_segment_width_parameters(initial_solution, options=options)
        solution = initial_solution
        if max_error is not None and max_error < tol:  # the warm start is already good enough
            self.iter_info[self.iter_count] = max_error
            self._say(f"Solved to acceptable tolerance {tol}", max_error)
            max_iter = 0
        for it in range(max_iter):
            self._nlp_sw_params = widths_next
            if self.iter_count > 0:
                self.iter_info[self.iter_count] = max_error
                if self.iter_count > 4:  # stagnation of the max residual over the last four passes
                    recent = np.mean(list(self.iter_info.values())[-4:])
                    if abs(max_error - recent) < 0.05 * abs(max_error):
                        self._say("Stopping the iterations: Change in max error is < 5%")
                        self._nlp_sw_params = widths_prev
                        break
            if it > 0:
                new, old = np.asarray(self._nlp_sw_params, float), np.asarray(widths_prev, float)
                if (np.abs(new - old) / new <= self._TOL_SEG_WIDTH_CHANGE).all():
                    self._say("Stopping the iterations: Change in width less than 5%", max_error)
                    self._nlp_sw_params = widths_prev
                    break
            if it == 0 and initial_solution is None:
                max_error = None
            if max_error is not None:
                self._say(f"Iteration : {it}, {max_error}")
            inputs = self.get_solver_warm_start_input_parameters(solution)
            inputs["p"] = self._nlp_sw_params
            solution = self.nlp_solver(**inputs, **self.nlp_bounds)  # same plan, new widths
            widths_prev = np.array(self._nlp_sw_params, dtype=float)
            widths_next, max_error = self.get_segment_width_parameters(solution, options=options)
            self.iter_count += 1
            if max_error is not None and max_error < tol:
                self.iter_info[self.iter_count] = max_error
                self._say(f"Solved to acceptable tolerance {tol}", max_error)
                break
            if it == max_iter - 1:
                self.iter_info[self.iter_count] = max_error
                self._say("Stopping the iterations: Iteration limit exceeded")
        self._say(f"H-Adaptive Iter., max_residual : {self.iter_count}, {max_error}")
        if solution is not None:
            self._say(" Optimal cost (J): ", solution["f"], "\n")
        self._say(f" Solved in {round((time.monotonic() - tic) * 1e3, 3)} ms\n")
        return solution

    # ------------------------------------------------------------------ width updates
    def get_segment_width_parameters(self, solution, options={"method": "residual", "sub_method": "merge_split"}):
        """(widths of all phases, max residual or None) from a solution (mpopt.py:2474-2522)."""
        K, P = self.n_segments, self._ocp.n_phases
        equal = [1 / K] * (K * P)
        if K == 1 or solution is None:
            return equal, None
        if not hasattr(self, "_nlp_sw_params"):
            self._nlp_sw_params = equal
        method = options.get("method")
        if method == "control_slope":
            return self.compute_seg_width_based_on_input_slope(solution)
        if method == "residual":
            return self.compute_seg_width_based_on_residuals(solution, method=options.get("sub_method", "equal_area"))
        return equal, None

    def _phase_widths(self, phase):
        K = self.n_segments
        return np.asarray(self._nlp_sw_params, dtype=float)[K * phase: K * (phase + 1)]

    def compute_seg_width_based_on_residuals(self, solution, method: str = "merge_split"):
        """mpopt.py:2524-2592."""
        _, residuals = self.get_dynamics_residuals(solution)  # GPU: mpx_eval_residuals
        widths, max_error = [], 0
        for phase in range(self._ocp.n_phases):
            peak = max(_seg_peak(r) for r in residuals[phase])
            max_error = max(max_error, peak)
            old = self._phase_widths(phase)
            if peak < self.tol_residual[phase]:
                self._say(f"Solved phase {phase} to acceptable tolerance {self.tol_residual[phase]}")
                widths.append(old)
                continue
            new = self.refine_segment_widths_based_on_residuals(residuals[phase], old, ERR_TOL=self.tol_residual[phase],
                                                                method=method)
            if method == "equal_area":
                new = 0.4 * np.asarray(new, float) + 0.6 * old  # damped update (:2588-2591)
            widths.append(np.asarray(new, float))
        return np.concatenate(widths), max_error

    def refine_segment_widths_based_on_residuals(self, residuals, segment_widths, ERR_TOL: float = 1e-3,
                                                 method: str = "merge_split"):
        """mpopt.py:2594-2635."""
        if method == "merge_split":
            return self.merge_split_segments_based_on_residuals([_seg_peak(r) for r in residuals], segment_widths,
                                                                ERR_TOL=ERR_TOL)
        if method == "equal_area":
            profile = np.concatenate([np.linalg.norm(np.asarray(r, float), 2, axis=1) if r is not None else [0]
                                      for r in residuals])
            return self.get_roots_wrt_equal_area(profile, self.n_segments)
        return segment_widths

    @staticmethod
    def get_roots_wrt_equal_area(residuals, n_segments):
        """Widths that split the area under the (piecewise-linear) residual profile into equal parts
        (mpopt.py:2637-2660)."""
        r = np.asarray(residuals, dtype=float)
        n = len(r)
        cum = np.append(0, np.cumsum(0.5 * (r[:-1] + r[1:])))
        cum = cum / cum[-1]
        edges = np.zeros(n_segments + 1)
        for i in range(n_segments):
            target = (i + 1) / n_segments
            j = int((cum >= target).argmax())
            edges[i + 1] = (j - 1 + (target - cum[j - 1]) / (cum[j] - cum[j - 1])) / (n - 1)
        return [edges[i + 1] - edges[i] for i in range(n_segments)]

    @staticmethod
    def merge_split_segments_based_on_residuals(max_residuals, segment_widths, ERR_TOL: float = 1e-3):
        """Merge runs of segments below tolerance, hand the freed segments to the runs above it
        (mpopt.py:2663-2707)."""
        ns = len(segment_widths)
        ok = [max_residuals[k] < ERR_TOL for k in range(ns)]
        runs = [(flag, [k for k, _ in grp]) for flag, grp in itertools.groupby(enumerate(ok), key=lambda kv: kv[1])]
        n_bad = sum(1 for flag, _ in runs if not flag)
        if len(runs) == ns or n_bad == 0:  # nothing to merge, or nothing to split
            return segment_widths
        run_width = [sum(segment_widths[k] for k in ks) for _, ks in runs]
        n_free = ns - len(runs)
        share = [1 + int(n_free / n_bad)] * n_bad
        share[-1] += int(np.mod(n_free, n_bad))
        out, bad = [], 0
        for (flag, _), wsum in zip(runs, run_width):
            if flag:
                out.append(wsum)
            else:
                out += [wsum / share[bad]] * share[bad]
                bad += 1
        return np.array(out)

    def compute_seg_width_based_on_input_slope(self, solution):
        """Segment boundaries at the largest control slopes (mpopt.py:2709-2824)."""
        _, residuals = self.get_dynamics_residuals(solution)
        if not self._collocation_approximation_computed:
            self.compute_numerical_approximation()
        post = self.process_results(solution, plot=False, scaling=True)
        widths, max_error = [], 0.0
        for phase in range(self._ocp.n_phases):
            peak = max(_seg_peak(r) for r in residuals[phase])
            max_error = max(max_error, peak)
            old = self._phase_widths(phase)
            if peak < self.tol_residual[phase]:
                self._say(f"Solved phase {phase} to acceptable level {self.tol_residual[phase]}, residual: {peak}")
                widths.append(old)
                continue
            _, u, t, _ = post.get_trajectories(phase)  # scaled controls, time in the OCP's units
            t = np.asarray(t, float).reshape(-1)
            t0, tf = t[0], t[-1]
            taus = self.get_residual_grid_taus(phase)
            grid = self.get_interpolated_time_grid(t, taus, self.poly_orders, self.tau0, self.tau1)
            slope = np.abs(self._compD @ u)
            times = self.compute_time_at_max_values(grid[1:-1], t, slope, threshold=self._THRESHOLD_SLOPE)
            if len(times) == 0:
                widths.append(old)
                continue
            new = self.compute_segment_widths_at_times(times, self.n_segments, t0, tf)
            new = np.clip(new, self.lbh[phase], self.ubh[phase])
            widths.append(new / new.sum())
        return np.concatenate(widths), max_error

    @staticmethod
    def compute_time_at_max_values(t_grid, t_orig, du_orig, threshold: float = 0):
        """Interior node times whose control-slope 2-norm reaches the threshold, ordered by increasing slope
        (mpopt.py:2826-2852)."""
        mag = np.linalg.norm(np.asarray(du_orig, float), 2, axis=1)[1:-1]
        tt = np.asarray(t_orig, float).reshape(-1)[1:-1]
        keep = mag >= threshold
        return tt[keep][np.argsort(mag[keep], kind="stable")]

    @staticmethod
    def compute_segment_widths_at_times(times, n_segments, t0, tf):
        """Width fractions with segment boundaries at the given times (mpopt.py:2854-2874)."""
        times = np.array(times, dtype=float)
        n_avail = len(times)
        w = np.empty(n_segments)
        if n_avail > n_segments - 2:
            cuts = np.sort(times[:n_segments])
            w[0] = cuts[0] - t0
            w[1:n_segments - 1] = np.diff(cuts)[:n_segments - 2]
            w[n_segments - 1] = tf - cuts[n_segments - 2]
        else:
            cuts = np.sort(times)
            head, tail = cuts[0] - t0, tf - cuts[-1]
            n_req = n_segments - (n_avail - 1)
            n_head = 1 if n_req == 2 else 1 + int(head / (head + tail) * (n_req - 1))
            n_tail = n_req - n_head
            w[:n_head] = head / n_head
            w[n_head:n_head + n_avail - 1] = np.diff(cuts)
            w[n_head + n_avail - 1:] = tail / n_tail
        return w / (tf - t0)


class mpopt_adaptive(mpopt):
    """``mp.mpopt_adaptive``: the segment widths are decision variables of the NLP and are solved for together with the
    trajectory (/root/reference/mpopt/mpopt.py:2877-3375).

    Per phase ``z = [X, U, t0, tf, a, w_0 .. w_{K-1}]`` (:2938-2945), no NLP parameters (:3190-3191), rows
    ``[F, C, DU, TC, SW]`` (:3169) with ``SW = [sum(w) - 1, compI.U, compI.X, mid-point residuals]`` (:3034-3136).
    The same kernels evaluate ``F / C / DU / TC`` (they read the widths out of z); ``mpx_adapt_kernel`` adds the
    ``SW`` block and every ``d/dw`` entry.  Set ``mid_residuals``, ``lbh``, ``ubh``, ``tol_residual`` before the first
    ``solve`` / ``create_solver`` like in the reference (tests/test_mpopt.py:474)."""

    _SEG_WIDTH_MIN = 1e-4
    _SEG_WIDTH_MAX = 1.0
    _TOL_RESIDUAL = 1e-3

    def __init__(self, problem, n_segments: int = 1, poly_orders=[9], scheme: str = "LGR", **kwargs):
        super().__init__(problem, n_segments=n_segments, poly_orders=poly_orders, scheme=scheme, **kwargs)
        P = self._ocp.n_phases
        self.mid_residuals = True
        self.lbh = [self._SEG_WIDTH_MIN] * P
        self.ubh = [self._SEG_WIDTH_MAX] * P
        self.tol_residual = [self._TOL_RESIDUAL] * P

    @property
    def transcription(self):
        if self._tr is None:
            from .collocation import CollocationRoots
            from .nlp import Transcription

            tr = Transcription(self._ocp, self.n_segments, self.poly_orders, self.colloc_scheme,
                               tau_min=float(CollocationRoots._TAU_MIN), tau_max=float(CollocationRoots._TAU_MAX),
                               device=self.device, adaptive=True, mid_residuals=self.mid_residuals)
            tr.lbh, tr.ubh, tr.tol_residual = list(self.lbh), list(self.ubh), list(self.tol_residual)
            self._tr = tr
            o = self._ocp
            self._optimization_vars_per_phase = self._Npoints * (o.nx + o.nu) + o.na + 2 + self.n_segments
            self._variables_created = True
        return self._tr

    def _solution_widths(self, z):
        """The widths are decision variables of this NLP: the trajectories of a solution are laid out on ITS widths
        (mpopt.py:3248-3273 evaluates the time grid from Z and ignores the parameter), not on those of the last solve."""
        L, K = self.transcription.layout, self.n_segments
        z = np.asarray(z, dtype=float).reshape(-1)
        return np.concatenate([z[L.colW(ph, 0): L.colW(ph, 0) + K] for ph in range(self._ocp.n_phases)])

    def get_nlp_constrains_for_segment_widths(self, phase: int = 0):
        """(SW, SWmin, SWmax): row indices and bounds of the width block of one phase (mpopt.py:3034-3136)."""
        tr = self.transcription
        _, _, gmin, gmax = tr.bounds()
        a = tr.layout.phases[phase].gSW
        b = tr.layout.phases[phase + 1].gF if phase + 1 < len(tr.layout.phases) else tr.layout.g_events
        return np.arange(a, b), gmin[a:b], gmax[a:b]

    def get_segment_width_parameters(self, solution=None):
        return np.zeros(0)  # the NLP has no parameters (:3190-3191)

    def create_solver(self, solver: str = "ipopt", options={}):
        super().create_solver(solver=solver, options=options)

    def solve(self, initial_solution=None, reinitialize_nlp=False, solver="ipopt", nlp_solver_options={},
              mpopt_options={}, **kwargs):
        """mpopt.py:3207-3246: solve, then read the optimal width fractions out of x."""
        if (not self._nlpsolver_initialized) or reinitialize_nlp:
            self.create_solver(solver=solver, options=nlp_solver_options)
        inputs = self.get_solver_warm_start_input_parameters(initial_solution)
        solution = self.nlp_solver(**inputs, **self.nlp_bounds)
        L, K = self.transcription.layout, self.n_segments
        x = np.asarray(solution["x"], dtype=float).reshape(-1)
        self._nlp_sw_params = np.concatenate([x[L.colW(ph, 0): L.colW(ph, 0) + K] for ph in range(self._ocp.n_phases)])
        if not self._MUTE_:
            print(f"Optimal segment width fractions: {self._nlp_sw_params}")
        return solution
