"""Host side of the transcribed NLP: traces the OCP, creates the device plan, exposes the evaluators.

``Transcription`` is what ``mpopt.create_nlp`` + ``ca.nlpsol`` amount to in the reference
(/root/reference/mpopt/mpopt.py:574-639, :725-758): it fixes the variable / constraint layout,
the bounds (:546-570 and the ``*min/*max`` vectors of :234-519), the initial guess (:641-708) and
owns the four evaluators ``f, grad_f, g, jac_g`` -- here CUDA kernels behind the C ABI of
``include/mpx.h`` instead of CasADi's SX virtual machine.
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np

from . import _lib
from .layout import Layout
from .program import Program


def _u8(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8).reshape(-1))


class Transcription:
    def __init__(self, ocp, n_segments=1, poly_orders=9, scheme="LGR", tau_min=-1.0, tau_max=1.0, device=0,
                 drop_exact_zeros=True, segments=None, program=None, adaptive=False, mid_residuals=True,
                 width_bounds=(1e-4, 1.0), tol_residual=1e-3):
        """``adaptive``: transcribe the NLP of the reference's ``mpopt_adaptive`` (mpopt.py:2877-3375) instead -- the
        segment widths become decision variables appended to every phase of z, ``n_p = 0``, rows per phase
        ``[F, C, DU, TC, SW]``; ``mid_residuals``, ``width_bounds = (lbh, ubh)`` and ``tol_residual`` are that class's
        ``mid_residuals`` flag, ``lbh / ubh`` and ``tol_residual`` (:2896-2923)."""
        if scheme not in _lib.SCHEMES:
            raise ValueError(f"scheme must be one of {sorted(_lib.SCHEMES)} (got {scheme!r})")
        self.ocp = copy.deepcopy(ocp)  # the reference snapshots the OCP too (mpopt.py:77)
        self.K = int(n_segments)
        self.poly_orders = [int(poly_orders)] * self.K if isinstance(poly_orders, (int, np.integer)) else \
            [int(v) for v in poly_orders]
        if len(self.poly_orders) != self.K:
            raise AssertionError("len(poly_orders) must equal n_segments")  # mpopt.py:83
        self.scheme, self.tau0, self.tau1 = scheme, float(tau_min), float(tau_max)
        o = self.ocp
        self.nx, self.nu, self.na, self.P = o.nx, o.nu, o.na, o.n_phases
        self.N = sum(self.poly_orders) + 1  # mpopt.py:84
        self.program = program if program is not None else Program(o)
        self.device = int(device)
        self.adaptive, self.mid_residuals = bool(adaptive), bool(adaptive) and bool(mid_residuals)
        self.lbh, self.ubh = [float(width_bounds[0])] * self.P, [float(width_bounds[1])] * self.P
        self.tol_residual = [float(tol_residual)] * self.P

        # ---- description handed over the C ABI (arrays kept alive on self)
        L = _lib.lib()
        self._keep = []
        phases = (_lib.PhaseDesc * self.P)()
        self.has_mU = []
        for ph, pp in enumerate(self.program.phases):
            mid = bool(o.midu[ph]) and bool((np.asarray(o.lbu[ph]) > -np.inf).any() or
                                            (np.asarray(o.ubu[ph]) < np.inf).any())  # mpopt.py:346, :363-365
            self.has_mU.append(mid and not self.adaptive)
            finite = lambda lo, hi: bool((np.asarray(lo) > -np.inf).any() or (np.asarray(hi) < np.inf).any())
            arrs = [_u8(pp.pat_f()), _u8(pp.f_nz), _u8(pp.f_t()), _u8(pp.pat_c()), _u8(pp.c_t()), _u8(pp.pat_tc()),
                    _u8(pp.pat_hw()), _u8(pp.pat_ht()), _u8(pp.pat_hf())]
            self._keep += arrs
            d = phases[ph]
            d.n_path, d.n_term = pp.nc, pp.ntc
            (d.pat_f, d.f_nz, d.f_t, d.pat_c, d.c_t, d.pat_tc, d.pat_hw, d.pat_ht,
             d.pat_hf) = [_lib.ptr(a, _lib.c_u8p) for a in arrs]
            d.phi_nz = int(bool(pp.L_nz) or any(pp.f_nz))
            d.diff_u, d.midu = int(bool(o.diff_u[ph])), int(mid and not self.adaptive)
            d.du_continuity = int(bool(o.du_continuity[ph]) and not self.adaptive)
            d.sw_u = int(self.adaptive and finite(o.lbu[ph], o.ubu[ph]))  # mpopt.py:3066-3068
            d.sw_x = int(self.adaptive and finite(o.lbx[ph], o.ubx[ph]))  # mpopt.py:3075-3077
            d.cost_t = int(not pp.Lt.is_value(0.0))
        self.sw_u, self.sw_x = [bool(phases[ph].sw_u) for ph in range(self.P)], [bool(phases[ph].sw_x) for ph in range(self.P)]
        self.layout = Layout(self.program, self.poly_orders, [bool(v) for v in o.diff_u], self.has_mU,
                             [bool(v) and not self.adaptive for v in o.du_continuity],
                             len(o.phase_links) if self.P > 1 else 0,
                             adaptive=dict(sw_u=self.sw_u, sw_x=self.sw_x, mid_residuals=self.mid_residuals)
                             if self.adaptive else None)
        po = np.asarray(self.poly_orders, dtype=np.int32)
        sx, su, sa = (np.ascontiguousarray(np.asarray(v, dtype=float)) for v in (o.scale_x, o.scale_u, o.scale_a))
        links = np.asarray(o.phase_links if self.P > 1 else [], dtype=np.int32).reshape(-1)
        self._keep += [po, sx, su, sa, links, phases]
        desc = _lib.ProblemDesc()
        desc.nx, desc.nu, desc.na, desc.n_phases = self.nx, self.nu, self.na, self.P
        desc.phases = phases
        desc.n_segments, desc.poly_orders = self.K, _lib.ptr(po, _lib.c_i32p)
        desc.scheme, desc.tau_min, desc.tau_max = _lib.SCHEMES[scheme], self.tau0, self.tau1
        desc.scale_x, desc.scale_u, desc.scale_a = _lib.ptr(sx), _lib.ptr(su), _lib.ptr(sa)
        desc.scale_t = float(o.scale_t)
        desc.n_links, desc.links = len(links) // 2, _lib.ptr(links, _lib.c_i32p)
        desc.drop_exact_zeros = int(bool(drop_exact_zeros))
        desc.program_key = self.program.key().encode()
        desc.program_source = self.program.cuda_source().encode()
        desc.device = self.device
        desc.seg_begin, desc.seg_end = (0, 0) if segments is None else (int(segments[0]), int(segments[1]))
        desc.adaptive, desc.mid_residuals = int(self.adaptive), int(self.mid_residuals)
        self.segments = (0, self.K) if segments is None else (int(segments[0]), int(segments[1]))
        plan = C.c_void_p()
        _lib.check(L.mpx_plan_create(C.byref(desc), C.byref(plan)))
        self._plan, self._L = plan, L

        s = [C.c_int64() for _ in range(4)]
        _lib.check(L.mpx_sizes(plan, *[C.byref(v) for v in s]))
        self.n_z, self.n_p, self.n_g, self.nnz = (int(v.value) for v in s)
        self._structure = None

    def __del__(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan:
            self._L.mpx_plan_destroy(plan)

    # ------------------------------------------------------------------ structure / tables
    def structure(self):
        """(rowptr, colind) of jac_g, CSR, int64, sorted columns."""
        if self._structure is None:
            rp, ci = np.empty(self.n_g + 1, np.int64), np.empty(self.nnz, np.int64)
            _lib.check(self._L.mpx_jac_structure(self._plan, _lib.ptr(rp, _lib.c_i64p), _lib.ptr(ci, _lib.c_i64p)))
            self._structure = (rp, ci)
        return self._structure

    def structure_ccs(self):
        """(colptr, rowind, perm): CasADi's column-compressed order; ccs_values = csr_values[perm]."""
        cp, ri, pm = np.empty(self.n_z + 1, np.int64), np.empty(self.nnz, np.int64), np.empty(self.nnz, np.int64)
        _lib.check(self._L.mpx_jac_structure_ccs(self._plan, *[_lib.ptr(a, _lib.c_i64p) for a in (cp, ri, pm)]))
        return cp, ri, pm

    def tables(self, deg):
        n1 = deg + 1
        r, D, w, Cm = np.empty(n1), np.empty((n1, n1)), np.empty(n1), np.empty((deg, n1))
        _lib.check(self._L.mpx_plan_tables(self._plan, deg, *[_lib.ptr(a) for a in (r, D, w, Cm)]))
        return r, D, w, Cm

    def shard_runs(self, kind):
        n = C.c_int64()
        _lib.check(self._L.mpx_shard_runs(self._plan, kind, None, C.byref(n)))
        runs = np.empty(2 * n.value, np.int64)
        _lib.check(self._L.mpx_shard_runs(self._plan, kind, _lib.ptr(runs, _lib.c_i64p), C.byref(n)))
        return runs.reshape(-1, 2)

    # ------------------------------------------------------------------ evaluators (host buffers)
    def _zp(self, z, p):
        z = np.ascontiguousarray(z, dtype=float)
        p = self.seg_width_params() if (p is None or self.adaptive) else np.ascontiguousarray(p, dtype=float)
        if z.shape != (self.n_z,) or p.shape != (self.n_p,):
            raise ValueError(f"expected z of shape ({self.n_z},) and p of shape ({self.n_p},)")
        return z, p

    def f(self, z, p=None):
        z, p = self._zp(z, p)
        out = np.empty(1)
        _lib.check(self._L.mpx_eval_f(self._plan, _lib.ptr(z), _lib.ptr(p), _lib.ptr(out)))
        return float(out[0])

    def grad_f(self, z, p=None, out=None):
        z, p = self._zp(z, p)
        out = np.empty(self.n_z) if out is None else out
        _lib.check(self._L.mpx_eval_grad_f(self._plan, _lib.ptr(z), _lib.ptr(p), None, _lib.ptr(out)))
        return out

    def g(self, z, p=None, out=None):
        z, p = self._zp(z, p)
        out = np.empty(self.n_g) if out is None else out
        _lib.check(self._L.mpx_eval_g(self._plan, _lib.ptr(z), _lib.ptr(p), _lib.ptr(out)))
        return out

    def jac_g_values(self, z, p=None, out=None, g_out=None):
        z, p = self._zp(z, p)
        out = np.empty(self.nnz) if out is None else out
        _lib.check(self._L.mpx_eval_jac_g(self._plan, _lib.ptr(z), _lib.ptr(p), _lib.ptr(g_out), _lib.ptr(out)))
        return out

    # ---- the host hop (include/mpx.h): registered caller buffers, dynamic fetch
    def host_register(self, array):
        """Pin + map a caller-owned numpy array for the life of the plan (or until ``host_unregister``)."""
        _lib.check(self._L.mpx_host_register(self._plan, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        _lib.check(self._L.mpx_host_unregister(self._plan, array.ctypes.data))

    def jac_g_values_dynamic(self, z, p=None, out=None, g_out=None):
        """Like ``jac_g_values`` into a registered ``out``: after the first call only the z- / p-dependent entries are
        rewritten (``out`` must not be modified in between)."""
        z, p = self._zp(z, p)
        _lib.check(self._L.mpx_eval_jac_g_dynamic(self._plan, _lib.ptr(z), _lib.ptr(p), _lib.ptr(g_out), _lib.ptr(out)))
        return out

    def dynamic_positions(self):
        """CSR positions (ascending, int32) of the Jacobian entries that depend on z or p."""
        n = C.c_int64()
        _lib.check(self._L.mpx_jac_dynamic_count(self._plan, C.byref(n)))
        pos = np.empty(int(n.value), np.int32)
        if pos.size:
            _lib.check(self._L.mpx_jac_dynamic_positions(self._plan, _lib.ptr(pos, _lib.c_i32p)))
        return pos

    def jac_g_packed(self, z, p=None, out=None, g_out=None):
        """The dynamic entries only, packed in the order of ``dynamic_positions()``."""
        z, p = self._zp(z, p)
        out = np.empty(len(self.dynamic_positions())) if out is None else out
        _lib.check(self._L.mpx_eval_jac_g_packed(self._plan, _lib.ptr(z), _lib.ptr(p), _lib.ptr(g_out), _lib.ptr(out)))
        return out

    def jac_g(self, z, p=None):
        import scipy.sparse as sp

        rp, ci = self.structure()
        return sp.csr_matrix((self.jac_g_values(z, p), ci, rp), shape=(self.n_g, self.n_z))

    # ------------------------------------------------------------------ Hessian of the Lagrangian
    def hess_structure(self):
        """(rowptr, colind) of the lower triangle of the Lagrangian Hessian, CSR, int64, sorted columns."""
        if getattr(self, "_hstructure", None) is None:
            n = C.c_int64()
            _lib.check(self._L.mpx_hess_structure(self._plan, C.byref(n), None, None))
            rp, ci = np.empty(self.n_z + 1, np.int64), np.empty(n.value, np.int64)
            _lib.check(self._L.mpx_hess_structure(self._plan, None, _lib.ptr(rp, _lib.c_i64p), _lib.ptr(ci, _lib.c_i64p)))
            self._hstructure = (rp, ci)
        return self._hstructure

    def hess_l_values(self, z, p=None, lam_f=1.0, lam_g=None, out=None):
        z, p = self._zp(z, p)
        lam = np.zeros(self.n_g) if lam_g is None else np.ascontiguousarray(lam_g, dtype=float)
        if lam.shape != (self.n_g,):
            raise ValueError(f"expected lam_g of shape ({self.n_g},)")
        nnz = len(self.hess_structure()[1])
        out = np.empty(nnz) if out is None else out
        _lib.check(self._L.mpx_eval_hess_l(self._plan, _lib.ptr(z), _lib.ptr(p), float(lam_f), _lib.ptr(lam), _lib.ptr(out)))
        return out

    @property
    def hess_zero_fill(self):
        """Adaptive plans: -1 before the first Hessian evaluation, 0 when it runs without zero-filling its output (every
        entry of the pattern has a writer; probed once on a NaN-filled buffer), 1 when it zero-fills."""
        st = C.c_int(0)
        _lib.check(self._L.mpx_hess_zero_fill(self._plan, C.byref(st)))
        return st.value

    def hess_l(self, z, p=None, lam_f=1.0, lam_g=None):
        """Lower triangle of the Hessian of ``lam_f * f + lam_g . g`` as scipy.sparse.csr_matrix (CasADi's nlp_hess_l)."""
        import scipy.sparse as sp

        rp, ci = self.hess_structure()
        return sp.csr_matrix((self.hess_l_values(z, p, lam_f, lam_g), ci, rp), shape=(self.n_z, self.n_z))

    # ------------------------------------------------------------------ interpolation / residuals
    def residuals(self, z, p=None, phase=0, taus=None, derivatives=True):
        """Interpolate the solution ``z`` and evaluate the dynamics residual at per-segment local abscissae.

        ``taus``: one array per segment (possibly empty) of points in ``[tau_min, tau_max]`` -- the reference's
        ``target_nodes`` (mpopt.py:1489-1542, :1428-1487).  Returns a dict of arrays whose rows are the points, segment by
        segment: ``xi`` (n, nx), ``ui`` (n, nu), ``ti`` (n,), and with ``derivatives`` also ``dxi``, ``dui`` and
        ``res = dxi - h Sx f``; ``counts`` is the number of points per segment."""
        z, p = self._zp(z, p)
        seg, tau, counts = self._pack_points(taus)
        n = len(tau)
        out = {"xi": np.empty((n, self.nx)), "ui": np.empty((n, self.nu)), "ti": np.empty(n), "counts": counts}
        if derivatives:
            out.update(dxi=np.empty((n, self.nx)), dui=np.empty((n, self.nu)), res=np.empty((n, self.nx)))
        ptr = lambda k: _lib.ptr(out[k]) if k in out and out[k].size else None
        _lib.check(self._L.mpx_eval_residuals(self._plan, _lib.ptr(z), _lib.ptr(p), int(phase), n,
                                              _lib.ptr(seg, _lib.c_i32p), _lib.ptr(tau), ptr("xi"), ptr("ui"), ptr("ti"),
                                              ptr("dxi"), ptr("dui"), ptr("res")))
        return out

    def _pack_points(self, taus):
        """(segment of each point, local abscissa of each point, points per segment) from the reference's per-segment
        lists; a 2-D array [K, m] (the same number of points in every segment: mid points, uniform grids) is packed
        without a Python loop over the segments."""
        if isinstance(taus, np.ndarray) and taus.ndim == 2:
            if taus.shape[0] != self.K:
                raise ValueError("taus must hold one row per segment")
            m = taus.shape[1]
            key = (self.K, m)
            if getattr(self, "_seg_cache", (None,))[0] != key:
                self._seg_cache = (key, np.repeat(np.arange(self.K, dtype=np.int32), m), [m] * self.K)
            return self._seg_cache[1], np.ascontiguousarray(taus, dtype=float).reshape(-1), self._seg_cache[2]
        if taus is None or len(taus) != self.K:
            raise ValueError("taus must hold one array per segment")
        counts = np.fromiter(map(len, taus), dtype=np.int64, count=self.K)
        n = int(counts.sum())
        seg = np.repeat(np.arange(self.K, dtype=np.int32), counts)
        tau = np.ascontiguousarray(np.concatenate(taus), dtype=float).reshape(-1) if n else np.zeros(0)
        return seg, tau, counts.tolist()

    def state_residuals(self, z, p=None, phase=0, taus=None):
        """State residual by quadrature at per-segment target points (mpopt.py:989-1076): ``xint`` = state at the
        segment start + integral of the interpolated ``h Sx f`` up to each point, ``res_x = xi - xint``; also ``ui``,
        ``ti``, ``counts``.  All arrays have one row per point, segment by segment."""
        z, p = self._zp(z, p)
        if taus is None or len(taus) != self.K:
            raise ValueError("taus must hold one array per segment")
        counts = np.fromiter(map(len, taus), dtype=np.int64, count=self.K)
        n = int(counts.sum())
        seg = np.repeat(np.arange(self.K, dtype=np.int32), counts)
        tau = np.ascontiguousarray(np.concatenate(taus), dtype=float).reshape(-1) if n else np.zeros(0)
        out = {"xint": np.empty((n, self.nx)), "res_x": np.empty((n, self.nx)), "ui": np.empty((n, self.nu)),
               "ti": np.empty(n), "counts": counts.tolist()}
        if n:
            _lib.check(self._L.mpx_eval_state_residuals(self._plan, _lib.ptr(z), _lib.ptr(p), int(phase), n,
                                                        _lib.ptr(seg, _lib.c_i32p), _lib.ptr(tau), _lib.ptr(out["xint"]),
                                                        _lib.ptr(out["ui"]) if self.nu else None, _lib.ptr(out["ti"]),
                                                        _lib.ptr(out["res_x"])))
        return out

    def second_derivatives(self, z, p=None, phase=0, taus=None):
        """d2/dtau2 of the state / control interpolants at per-segment local abscissae (mpopt.py:1285-1358: composite
        ``get_diff_matrix(order=2)`` times X, U).  Returns ``ti`` (n,), ``ddxi`` (n, nx), ``ddui`` (n, nu), ``counts``."""
        z, p = self._zp(z, p)
        if taus is None or len(taus) != self.K:
            raise ValueError("taus must hold one array per segment")
        counts = np.fromiter(map(len, taus), dtype=np.int64, count=self.K)
        n = int(counts.sum())
        seg = np.repeat(np.arange(self.K, dtype=np.int32), counts)
        tau = np.ascontiguousarray(np.concatenate(taus), dtype=float).reshape(-1) if n else np.zeros(0)
        out = {"ti": np.empty(n), "ddxi": np.empty((n, self.nx)), "ddui": np.empty((n, self.nu)), "counts": counts.tolist()}
        if n:
            _lib.check(self._L.mpx_eval_second_derivatives(self._plan, _lib.ptr(z), _lib.ptr(p), int(phase), n,
                                                           _lib.ptr(seg, _lib.c_i32p), _lib.ptr(tau), _lib.ptr(out["ti"]),
                                                           _lib.ptr(out["ddxi"]) if self.nx else None,
                                                           _lib.ptr(out["ddui"]) if self.nu else None))
        return out

    # ------------------------------------------------------------------ evaluators (device pointers)
    def g_jac_dev(self, z_ptr, p_ptr, g_ptr, vals_ptr, stream=None):
        _lib.check(self._L.mpx_eval_g_jac_dev(self._plan, z_ptr, p_ptr, g_ptr, vals_ptr, stream))

    def g_jac_dev_peers(self, z_ptr, p_ptr, g_ptr, vals_ptr, peer_g, peer_vals, stream=None):
        """Fused evaluation + all-gather: every store also goes to the peers' buffers (device pointers obtained with
        ``mpopt_b200.shard.PeerBuffers``)."""
        n = len(peer_g)
        pg = (C.c_void_p * max(n, 1))(*peer_g)
        pv = (C.c_void_p * max(n, 1))(*peer_vals)
        _lib.check(self._L.mpx_eval_g_jac_dev_peers(self._plan, z_ptr, p_ptr, g_ptr, vals_ptr, n, pg, pv, stream))

    def f_grad_dev(self, z_ptr, p_ptr, f_ptr, grad_ptr, stream=None):
        _lib.check(self._L.mpx_eval_f_grad_dev(self._plan, z_ptr, p_ptr, f_ptr, grad_ptr, stream))

    def sync(self):
        _lib.check(self._L.mpx_sync(self._plan))

    def trace(self):
        """Timeline records of the last g + jac_g launch (plan created under MPX_TRACE=1): uint64 [warps, slots]."""
        nw, ns = C.c_int64(), C.c_int64()
        _lib.check(self._L.mpx_trace_read(self._plan, C.byref(nw), C.byref(ns), None))
        out = np.zeros((int(nw.value), int(ns.value)), dtype=np.uint64)
        if out.size:
            _lib.check(self._L.mpx_trace_read(self._plan, None, None, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    @property
    def launches(self):
        return int(self._L.mpx_launch_count(self._plan))

    @property
    def program_origin(self):
        return self._L.mpx_program_origin(self._plan).decode()

    # ------------------------------------------------------------------ bounds, parameters, initial guess
    def seg_width_params(self):
        """Equal segment widths summing to 1 per phase (mpopt.py:710-723); empty for the adaptive NLP, whose widths
        are decision variables (:3190-3191)."""
        return np.zeros(0) if self.adaptive else np.full(self.K * self.P, 1.0 / self.K)

    def bounds(self):
        """(Zmin, Zmax, Gmin, Gmax) in the layout of z and g."""
        o, N, K, nx, nu = self.ocp, self.N, self.K, self.nx, self.nu
        zlo, zhi, glo, ghi = [], [], [], []
        for ph, pp in enumerate(self.program.phases):
            xlo = np.repeat((np.asarray(o.lbx[ph], float) * o.scale_x)[:, None], N, axis=1)
            xhi = np.repeat((np.asarray(o.ubx[ph], float) * o.scale_x)[:, None], N, axis=1)
            if ph == 0:  # initial state pinned in phase 0 only (mpopt.py:550-551)
                xlo[:, 0] = xhi[:, 0] = np.asarray(o.x00[0], float) * o.scale_x
            ulo = np.repeat(np.asarray(o.lbu[ph], float) * o.scale_u, N)
            uhi = np.repeat(np.asarray(o.ubu[ph], float) * o.scale_u, N)
            zlo += [xlo.reshape(-1), ulo, np.atleast_1d(o.lbt0[ph] * o.scale_t), np.atleast_1d(o.lbtf[ph] * o.scale_t),
                    np.asarray(o.lba[ph], float) * o.scale_a]
            zhi += [xhi.reshape(-1), uhi, np.atleast_1d(o.ubt0[ph] * o.scale_t), np.atleast_1d(o.ubtf[ph] * o.scale_t),
                    np.asarray(o.uba[ph], float) * o.scale_a]
            if self.adaptive:  # width variables (mpopt.py:2958, :2976)
                zlo.append(np.full(K, self.lbh[ph]))
                zhi.append(np.full(K, self.ubh[ph]))
            glo += [np.full(nx * N, float(o.LB_DYNAMICS)), np.full(pp.nc * N, float(o.LB_PATH_CONSTRAINTS))]
            ghi += [np.full(nx * N, float(o.UB_DYNAMICS)), np.full(pp.nc * N, float(o.UB_PATH_CONSTRAINTS))]
            if o.diff_u[ph]:
                glo.append(np.full(nu * N, float(o.lbdu[ph])))
                ghi.append(np.full(nu * N, float(o.ubdu[ph])))
            if self.has_mU[ph] and not self.adaptive:
                glo.append(np.repeat(np.asarray(o.lbu[ph], float) * o.scale_u, N - 1))
                ghi.append(np.repeat(np.asarray(o.ubu[ph], float) * o.scale_u, N - 1))
            if o.du_continuity[ph] and K > 1 and not self.adaptive:
                glo.append(np.zeros(nu * (K - 1)))
                ghi.append(np.zeros(nu * (K - 1)))
            glo.append(np.full(pp.ntc, float(o.LB_TERMINAL_CONSTRAINTS)))
            ghi.append(np.full(pp.ntc, float(o.UB_TERMINAL_CONSTRAINTS)))
            if self.adaptive:  # SW block (mpopt.py:3037-3130)
                glo.append(np.zeros(1)), ghi.append(np.zeros(1))
                if self.sw_u[ph]:
                    glo.append(np.repeat(np.asarray(o.lbu[ph], float) * o.scale_u, N - 1))
                    ghi.append(np.repeat(np.asarray(o.ubu[ph], float) * o.scale_u, N - 1))
                if self.sw_x[ph]:
                    glo.append(np.repeat(np.asarray(o.lbx[ph], float) * o.scale_x, N - 1))
                    ghi.append(np.repeat(np.asarray(o.ubx[ph], float) * o.scale_x, N - 1))
                if self.mid_residuals:
                    glo.append(np.full(nx * (N - 1), -self.tol_residual[ph]))
                    ghi.append(np.full(nx * (N - 1), self.tol_residual[ph]))
        if self.P > 1:
            n = len(o.phase_links)
            # the reference indexes lbe/ube by link ordinal, not by phase id (mpopt.py:491-497)
            glo += [np.concatenate([np.asarray(o.lbe[i], float) * o.scale_x for i in range(n)]), np.zeros(n * nu),
                    np.zeros(n)]
            ghi += [np.concatenate([np.asarray(o.ube[i], float) * o.scale_x for i in range(n)]), np.zeros(n * nu),
                    np.zeros(n)]
        cat = lambda parts: np.concatenate([np.asarray(a, float).reshape(-1) for a in parts])
        return cat(zlo), cat(zhi), cat(glo), cat(ghi)

    def initial_guess(self):
        """Linear interpolation between the OCP's start/end guesses (mpopt.py:641-708).

        States are laid out state-major like z; controls come out node-major exactly as in the
        reference (its quirk, SURVEY.md Q3 -- harmless unless nu > 1 and u00 != uf0)."""
        o, N = self.ocp, self.N
        parts = []
        for ph in range(self.P):
            x0, xf = np.asarray(o.x00[ph], float) * o.scale_x, np.asarray(o.xf0[ph], float) * o.scale_x
            u0, uf = np.asarray(o.u00[ph], float) * o.scale_u, np.asarray(o.uf0[ph], float) * o.scale_u
            ta, tb = float(np.ravel(o.t00[ph])[0]) * o.scale_t, float(np.ravel(o.tf0[ph])[0]) * o.scale_t
            ts = np.linspace(ta, tb, N)
            X = x0[None, :] + ((xf - x0) / (tb - ta))[None, :] * (ts - ta)[:, None]
            U = u0[None, :] + ((uf - u0) / (tb - ta))[None, :] * (ts - ta)[:, None]
            parts += [X.T.reshape(-1), U.reshape(-1), [ta], [tb], np.asarray(o.a0[ph], float) * o.scale_a]
            if self.adaptive:  # mpopt.py:3030
                parts.append(np.full(self.K, 1.0 / self.K))
        return np.concatenate([np.asarray(a, float).reshape(-1) for a in parts])
