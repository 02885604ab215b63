"""Host logic of the h-adaptive driver that needs no GPU: the width-update rules of mpopt_h_adaptive
(/root/reference/mpopt/mpopt.py:2637-2707, :2826-2874) on hand-checkable inputs."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def mp():
    from mpopt_b200 import mp as _mp

    return _mp


def test_h_adaptive_static_helpers(mp):
    H = mp.mpopt_h_adaptive
    # equal residual everywhere -> equal widths
    w = H.get_roots_wrt_equal_area(np.ones(21), 4)
    assert np.allclose(w, 0.25)
    # two good segments merge, the bad one is split in two
    w = H.merge_split_segments_based_on_residuals([1e-6, 1e-6, 1.0], [0.25, 0.25, 0.5], ERR_TOL=1e-3)
    assert np.allclose(w, [0.5, 0.25, 0.25])
    # nothing to merge: unchanged
    w0 = [0.5, 0.5]
    assert H.merge_split_segments_based_on_residuals([1.0, 1e-6], w0, ERR_TOL=1e-3) is w0
    w = H.compute_segment_widths_at_times(np.array([1.0, 3.0]), 3, 0.0, 4.0)
    assert np.allclose(w, [0.25, 0.5, 0.25])
    w = H.compute_segment_widths_at_times(np.array([2.0]), 4, 0.0, 4.0)
    assert abs(w.sum() - 1) < 1e-12 and len(w) == 4


def test_equal_area_moves_boundaries_towards_the_residual_peak(mp):
    """A residual profile concentrated in the last quarter pulls three of four segment boundaries into it."""
    r = np.concatenate([np.full(30, 1e-3), np.full(10, 1.0)])
    w = np.asarray(mp.mpopt_h_adaptive.get_roots_wrt_equal_area(r, 4))
    assert abs(w.sum() - 1) < 1e-12 and (w > 0).all()
    assert w[0] > 0.7 and w[1:].max() < 0.1


def test_time_at_max_values_orders_by_slope_and_applies_threshold(mp):
    t = np.linspace(0.0, 1.0, 6)
    du = np.array([[9.0], [0.05], [3.0], [1.0], [2.0], [9.0]])
    times = mp.mpopt_h_adaptive.compute_time_at_max_values(None, t, du, threshold=0.1)
    assert np.allclose(times, [0.6, 0.8, 0.4])   # interior nodes only, ascending slope, 0.05 filtered out


def test_adaptive_classes_keep_the_reference_knobs(mp):
    H, A = mp.mpopt_h_adaptive, mp.mpopt_adaptive
    assert (H._SEG_WIDTH_MIN, H._SEG_WIDTH_MAX, H._TOL_SEG_WIDTH_CHANGE, H._TOL_RESIDUAL) == (1e-5, 1, 0.05, 1e-2)
    assert (H._DEFAULT_METHOD, H._DEFAULT_SUB_METHOD, H._THRESHOLD_SLOPE) == ("residual", "equal_area", 1e-1)
    assert (A._SEG_WIDTH_MIN, A._SEG_WIDTH_MAX, A._TOL_RESIDUAL) == (1e-4, 1.0, 1e-3)


class _Scripted:
    """Stand-ins for the solver and the residual evaluation: the outer loop's control flow needs no GPU."""

    def __init__(self, mpo, peaks):
        self.mpo, self.peaks, self.calls, self.widths_seen = mpo, list(peaks), 0, []
        mpo._nlpsolver_initialized = True
        mpo.nlp_bounds = {}
        mpo.nlp_solver = self.solve
        mpo.get_solver_warm_start_input_parameters = lambda sol=None: {"x0": np.zeros(3)}
        mpo.get_dynamics_residuals = self.residuals

    def solve(self, x0=None, p=None, **kw):
        self.widths_seen.append(np.array(p, dtype=float))
        return {"x": np.zeros(3), "f": float(len(self.widths_seen))}

    def residuals(self, solution, **kw):
        peak = self.peaks[min(self.calls, len(self.peaks) - 1)]
        self.calls += 1
        K = self.mpo.n_segments
        # two points per segment, one state; the last segment carries the peak
        res = [np.full((2, 1), 1e-6) for _ in range(K - 1)] + [np.full((2, 1), peak)]
        return [[None] * K], [res]


def _h_adaptive(mp, K=4):
    from mpopt_b200.problems import moon_lander

    mp.mpopt._MUTE_ = True
    return mp.mpopt_h_adaptive(moon_lander(), K, 3)


def test_h_adaptive_loop_stops_at_the_residual_tolerance(mp):
    """mpopt.py:2393-2460: first pass with equal widths, refine while the max residual exceeds tol_residual, stop as
    soon as it does not; iter_info records the residual after every pass."""
    mpo = _h_adaptive(mp)
    s = _Scripted(mpo, peaks=[0.5, 0.2, 5e-3])
    sol = mpo.solve(max_iter=10, mpopt_options={"method": "residual", "sub_method": "merge_split"})
    assert mpo.iter_count == 3 and sol["f"] == 3.0
    assert np.allclose(s.widths_seen[0], 0.25)                      # equal widths first (mpopt.py:2474-2490)
    assert np.allclose(s.widths_seen[1], [0.75, 0.25 / 3, 0.25 / 3, 0.25 / 3])  # three good segments merged, the bad one split
    assert list(mpo.iter_info.values())[-1] == 5e-3 and abs(sum(s.widths_seen[2]) - 1) < 1e-12


def test_h_adaptive_loop_stops_when_the_widths_settle_or_iterations_run_out(mp):
    mpo = _h_adaptive(mp)
    s = _Scripted(mpo, peaks=[0.5] * 20)
    mpo.solve(max_iter=3, mpopt_options={"method": "residual", "sub_method": "merge_split"})
    assert mpo.iter_count == 3 and len(s.widths_seen) == 3           # iteration limit (:2455-2457)
    # a method that proposes the same widths again: "change in width less than 5 %" ends the loop after the second solve
    mpo2 = _h_adaptive(mp)
    s2 = _Scripted(mpo2, peaks=[0.5] * 20)
    mpo2.refine_segment_widths_based_on_residuals = lambda residuals, widths, ERR_TOL=1e-3, method="": widths
    mpo2.solve(max_iter=10, mpopt_options={"method": "residual", "sub_method": "keep"})
    assert len(s2.widths_seen) == 1 or len(s2.widths_seen) == 2
    assert np.allclose(mpo2._nlp_sw_params, 0.25)


def test_h_adaptive_single_segment_has_nothing_to_refine(mp):
    mpo = _h_adaptive(mp, K=1)
    w, err = mpo.get_segment_width_parameters({"x": np.zeros(3)}, options={"method": "residual"})
    assert w == [1.0] and err is None
