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
