"""Oracle (TEST INFRASTRUCTURE): interpolation of a solution and the dynamics residual at arbitrary points.

CPU restatement (numpy, float64) of the post-solve path the reference runs after every solve and inside the
h-adaptive loop, /root/reference/mpopt/mpopt.py:

  compute_interpolation_taus_corresponding_to_original_grid   :1205-1237
  get_residual_grid_taus                                      :1152-1203
  get_interpolated_time_grid                                  :1544-1573
  interpolate_single_phase                                    :1489-1542
  get_dynamics_residuals_single_phase                         :1428-1487
  get_dynamics_residuals ("relative" scaling)                 :1360-1426

Pinned by the reference's own known answers for the two static helpers (tests/test_mpopt.py:663-675 and :1161-1196);
the residual values themselves are unpinned in the reference (only magnitude bounds after a solve, tests:730-798).
"""
from __future__ import annotations

import numpy as np

from .collocation import mid_points
from .dual import Vec


def interpolation_taus_on_original_grid(nodes_req, seg_widths, tau0=0.0, tau1=1.0):
    """mpopt.py:1205-1237.  Global target nodes in [tau0, tau1] -> per-segment local taus in [tau0, tau1]; a node on a
    segment boundary belongs to the earlier segment, the very first node belongs to nobody."""
    nodes_req = np.asarray(nodes_req, dtype=float)
    csw = np.append(0, np.cumsum(seg_widths))
    assert abs(csw[-1] - 1) < 1e-6
    scaled = 0 + (1 - 0) / (tau1 - tau0) * (nodes_req - tau0)
    out = []
    for i, seg in enumerate(seg_widths):
        t = scaled[scaled > csw[i]]
        t = t[t <= csw[i + 1]]
        t = (t - csw[i]) / seg
        out.append(tau0 + (tau1 - tau0) / (1 - 0) * (t - 0))
    return out


def interpolated_time_grid(t_orig, taus, poly_orders, tau0, tau1):
    """mpopt.py:1544-1573: time of every interpolation point, from the node times of the original grid."""
    t_orig = np.asarray(t_orig, dtype=float)
    t_seg = [t_orig[0]] + [t_orig[sum(poly_orders[: i + 1])] for i in range(len(poly_orders))]
    return np.concatenate([t_seg[i] + (t_seg[i + 1] - t_seg[i]) * (0 + (1 - 0) / (tau1 - tau0) * (np.asarray(taus[i], float) - tau0))
                           for i in range(len(t_seg) - 1)])


def residual_grid_taus(ora, phase, grid_type, p=None, max_grid_points=15):
    """mpopt.py:1152-1203 (``_MAX_GRID_POINTS`` = 15, :53; found wrong (20) by running the reference, oracle/refrun)."""
    p = ora.seg_width_params() if p is None else np.asarray(p, dtype=float)
    if grid_type == "fixed":
        n_nodes = max(sum(ora.po) + 2, max_grid_points + 2)
        target = np.linspace(ora.tau0, ora.tau1, n_nodes)
        taus = interpolation_taus_on_original_grid(target, p[ora.K * phase: ora.K * (phase + 1)], ora.tau0, ora.tau1)
        taus[0] = taus[0][:-1]
        return taus
    if grid_type == "mid-points":
        return [mid_points(ora.tab.roots[d]) for d in ora.po]
    if grid_type == "spectral":
        from .collocation import roots

        r = roots(ora.scheme, max_grid_points + 2, ora.tau0, ora.tau1)[1:-1]
        return [np.array(r) for _ in ora.po]
    return None


def interpolate_phase(ora, z, p, phase, taus):
    """mpopt.py:1489-1542: (Xi, Ui, ti, DXi, DUi) at the per-segment local taus (rows = points, segment by segment)."""
    X, U, T0, TF, A = ora._unpack(phase, np.asarray(z, dtype=float))
    p = ora.seg_width_params() if p is None else np.asarray(p, dtype=float)
    _, t, _, _ = ora._time_grid(phase, T0, TF, ora._widths(phase, z, p)[0])
    CI = ora.tab.composite_interpolation(taus, 0)
    DI = ora.tab.composite_interpolation(taus, 1)
    ti = interpolated_time_grid(t, taus, ora.po, ora.tau0, ora.tau1)
    return CI @ X, CI @ U, ti, DI @ X, DI @ U


def states_from_dynamics_phase(ora, z, p, phase, taus):
    """mpopt.py:989-1076: per segment, Lagrange polynomials on the segment's own target points (:1025-1028), their
    integrals from tau0 to every target point (:1056-1058), x_int = x(segment start) + h_seg * quad^T (f * scale_x)
    (:1059-1061) and residual = x_I - x_int (:1062).  Returns flat (n_points, nx) arrays (xint, res_x) and ti."""
    from .collocation import quadrature_weights

    z = np.asarray(z, dtype=float)
    X, U, T0, TF, A = ora._unpack(phase, z)
    Xi, Ui, ti, DXi, _ = interpolate_phase(ora, z, p, phase, taus)
    _, res, F, n = dynamics_residuals_phase(ora, z, p, phase, taus)  # F = h_seg * Sx f at the points
    xint = np.zeros_like(Xi)
    off = np.concatenate([[0], np.cumsum(n)]).astype(int)
    for k in range(ora.K):
        r = np.asarray(taus[k], dtype=float)
        if len(r) == 0:
            continue
        xstart = X[ora.seg_start[k], :]
        Fk = F[off[k]: off[k + 1]]
        for i, tau in enumerate(r):
            w = quadrature_weights(r, ora.tau0, tau)
            xint[off[k] + i] = xstart + w @ Fk
    return xint, Xi - xint, ti


def second_derivatives_phase(ora, z, p, phase, taus):
    """mpopt.py:1285-1358: (ti, DDXi, DDUi) -- the composite second-order differentiation matrix at the local taus
    (get_composite_interpolation_Dmatrix_at(..., order=2)) applied to X and U."""
    X, U, T0, TF, A = ora._unpack(phase, np.asarray(z, dtype=float))
    p = ora.seg_width_params() if p is None else np.asarray(p, dtype=float)
    _, t, _, _ = ora._time_grid(phase, T0, TF, ora._widths(phase, z, p)[0])
    D2 = ora.tab.composite_interpolation(taus, 2)
    return interpolated_time_grid(t, taus, ora.po, ora.tau0, ora.tau1), D2 @ X, D2 @ U


def dynamics_residuals_phase(ora, z, p, phase, taus):
    """mpopt.py:1428-1487: residual = D_I X - h_seg * Sx f(Xi / Sx, Ui / Su, ti, a / Sa) at every point.
    Returns (ti, residual, F) as flat (n_points, nx) arrays plus the list of per-segment point counts."""
    o = ora.ocp
    z = np.asarray(z, dtype=float)
    p = ora.seg_width_params() if p is None else np.asarray(p, dtype=float)
    Xi, Ui, ti, DXi, _ = interpolate_phase(ora, z, p, phase, taus)
    _, _, T0, TF, A = ora._unpack(phase, z)
    t0, tf = T0 / o.scale_t, TF / o.scale_t
    n = [len(t) for t in taus]
    seg = np.repeat(np.arange(ora.K), n)
    w = p[ora.K * phase: ora.K * (phase + 1)]
    h = (tf - t0) / (ora.tau1 - ora.tau0) * w[seg]
    M = len(seg)
    if M == 0:
        return ti, np.zeros((0, ora.nx)), np.zeros((0, ora.nx)), n
    x = Vec(Xi[:, s] / o.scale_x[s] for s in range(ora.nx))
    u = Vec(Ui[:, c] / o.scale_u[c] for c in range(ora.nu))
    a = Vec(np.full(M, A[m] / o.scale_a[m]) for m in range(ora.na))
    f = o.get_dynamics(phase)(x, u, ti, a)  # one array (or constant) per state: not flattened element-wise
    f = list(f) if isinstance(f, (list, tuple)) else [f]
    assert len(f) == ora.nx
    F = np.stack([np.broadcast_to(np.asarray(fs, dtype=float), (M,)) * o.scale_x[s] for s, fs in enumerate(f)], axis=1)
    F = h[:, None] * F
    return ti, DXi - F, F, n


def dynamics_residuals(ora, z, p=None, nodes=None, grid_type="mid-points", residual_type=None):
    """mpopt.py:1360-1426: per phase, per segment lists (None for a segment without points)."""
    ti_all, res_all = [], []
    for ph in range(ora.P):
        taus = nodes[ph] if nodes is not None else residual_grid_taus(ora, ph, grid_type, p)
        ti, res, F, n = dynamics_residuals_phase(ora, z, p, ph, taus)
        off = np.concatenate([[0], np.cumsum(n)])
        res_seg = [res[off[k]: off[k + 1]] if n[k] else None for k in range(ora.K)]
        F_seg = [F[off[k]: off[k + 1]] if n[k] else None for k in range(ora.K)]
        ti_seg = [ti[off[k]: off[k + 1]] if n[k] else [] for k in range(ora.K)]
        if residual_type == "relative":  # global relative: divide by the largest |h Sx f| per state (:1400-1417)
            mx = np.zeros(ora.nx)
            for Fs in F_seg:
                if Fs is not None:
                    mx = np.maximum(mx, np.abs(Fs).max(axis=0))
            res_seg = [r / mx if r is not None else None for r in res_seg]
        ti_all.append(ti_seg), res_all.append(res_seg)
    return ti_all, res_all
