"""ctypes binding of the C ABI in include/mpx.h (libmpx.so, built in-tree by ``__graft_entry__.build``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a plan
is created, the caller gets an exception that says so.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpx.so")

MPX_OK, MPX_EINVAL, MPX_ENODEVICE, MPX_ECUDA, MPX_ENOPROGRAM, MPX_ELIMIT = 0, -1, -2, -3, -4, -5
SCHEMES = {"LGR": 0, "LGL": 1, "CGL": 2}

c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class PhaseDesc(C.Structure):
    _fields_ = [("n_path", C.c_int32), ("n_term", C.c_int32), ("pat_f", c_u8p), ("f_nz", c_u8p), ("f_t", c_u8p),
                ("pat_c", c_u8p), ("c_t", c_u8p), ("pat_tc", c_u8p), ("diff_u", C.c_int32), ("midu", C.c_int32),
                ("du_continuity", C.c_int32), ("cost_t", C.c_int32), ("pat_hw", c_u8p), ("pat_ht", c_u8p),
                ("sw_u", C.c_int32), ("sw_x", C.c_int32), ("pat_hf", c_u8p), ("phi_nz", C.c_int32)]


class ProblemDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("na", C.c_int32), ("n_phases", C.c_int32),
                ("phases", C.POINTER(PhaseDesc)), ("n_segments", C.c_int32), ("poly_orders", c_i32p),
                ("scheme", C.c_int32), ("tau_min", C.c_double), ("tau_max", C.c_double), ("scale_x", c_f64p),
                ("scale_u", c_f64p), ("scale_a", c_f64p), ("scale_t", C.c_double), ("n_links", C.c_int32),
                ("links", c_i32p), ("drop_exact_zeros", C.c_int32), ("program_key", C.c_char_p),
                ("program_source", C.c_char_p), ("device", C.c_int32), ("seg_begin", C.c_int32),
                ("seg_end", C.c_int32), ("adaptive", C.c_int32), ("mid_residuals", C.c_int32)]


class IpoptData(C.Structure):
    """mpx_ipopt_data: user_data of the IPOPT C-interface callbacks."""
    _fields_ = [("plan", C.c_void_p), ("p", c_f64p)]


MPX_STAGE_F, MPX_STAGE_GRAD, MPX_STAGE_G, MPX_STAGE_JAC, MPX_FETCH_JAC_CCS = 1, 2, 4, 8, 16

#: CasADi external-function symbols exported for each of these names (include/mpx.h, MPX_CASADI_DECLARE)
CASADI_FUNCTIONS = ("nlp_f", "nlp_g", "nlp_grad_f", "nlp_jac_g", "nlp_hess_l")
CASADI_SUFFIXES = ("", "_n_in", "_n_out", "_default_in", "_name_in", "_name_out", "_sparsity_in", "_sparsity_out",
                   "_work", "_alloc_mem", "_init_mem", "_free_mem", "_checkout", "_release", "_incref", "_decref")


class MpxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmpx error {code}: {msg}")
        self.code = code


#: every symbol include/mpx.h declares, with (restype, argtypes)
PROTOTYPES = {
    "mpx_version": (C.c_int, []),
    "mpx_last_error": (C.c_char_p, []),
    "mpx_collocation_tables": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, c_f64p, c_f64p,
                                         c_f64p, c_f64p]),
    "mpx_collocation_basis_at": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                           C.c_int32, c_f64p, c_f64p]),
    "mpx_collocation_weights": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double,
                                          C.c_double, c_f64p]),
    "mpx_plan_create": (C.c_int, [C.POINTER(ProblemDesc), C.POINTER(C.c_void_p)]),
    "mpx_plan_destroy": (None, [C.c_void_p]),
    "mpx_sizes": (C.c_int, [C.c_void_p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "mpx_jac_structure": (C.c_int, [C.c_void_p, c_i64p, c_i64p]),
    "mpx_jac_structure_ccs": (C.c_int, [C.c_void_p, c_i64p, c_i64p, c_i64p]),
    "mpx_plan_tables": (C.c_int, [C.c_void_p, C.c_int32, c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_shard_runs": (C.c_int, [C.c_void_p, C.c_int32, c_i64p, c_i64p]),
    "mpx_eval_f": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_grad_f": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_g": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_jac_g": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_f_grad_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpx_eval_g_jac_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpx_hess_structure": (C.c_int, [C.c_void_p, c_i64p, c_i64p, c_i64p]),
    "mpx_eval_hess_l": (C.c_int, [C.c_void_p, c_f64p, c_f64p, C.c_double, c_f64p, c_f64p]),
    "mpx_eval_hess_l_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpx_eval_residuals": (C.c_int, [C.c_void_p, c_f64p, c_f64p, C.c_int32, C.c_int64, c_i32p, c_f64p, c_f64p, c_f64p,
                                     c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_second_derivatives": (C.c_int, [C.c_void_p, c_f64p, c_f64p, C.c_int32, C.c_int64, c_i32p, c_f64p, c_f64p,
                                              c_f64p, c_f64p]),
    "mpx_eval_state_residuals": (C.c_int, [C.c_void_p, c_f64p, c_f64p, C.c_int32, C.c_int64, c_i32p, c_f64p, c_f64p,
                                           c_f64p, c_f64p, c_f64p]),
    "mpx_stage": (C.c_int, [C.c_void_p, c_f64p, c_f64p, C.c_int32]),
    "mpx_staged": (C.c_int, [C.c_void_p]),
    "mpx_fetch": (C.c_int, [C.c_void_p, C.c_int32, c_f64p]),
    "mpx_hess_l_staged": (C.c_int, [C.c_void_p, C.c_double, c_f64p, c_f64p]),
    "mpx_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mpx_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mpx_jac_dynamic_count": (C.c_int, [C.c_void_p, c_i64p]),
    "mpx_jac_dynamic_positions": (C.c_int, [C.c_void_p, c_i32p]),
    "mpx_eval_jac_g_dynamic": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_eval_jac_g_packed": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "mpx_gate": (C.c_int, [C.c_void_p, C.c_double]),
    "mpx_trace_read": (C.c_int, [C.c_void_p, c_i64p, c_i64p, C.POINTER(C.c_uint64)]),
    "mpx_hess_zero_fill": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mpx_ipopt_eval_f": (C.c_int, [C.c_int, c_f64p, C.c_int, c_f64p, C.c_void_p]),
    "mpx_ipopt_eval_grad_f": (C.c_int, [C.c_int, c_f64p, C.c_int, c_f64p, C.c_void_p]),
    "mpx_ipopt_eval_g": (C.c_int, [C.c_int, c_f64p, C.c_int, C.c_int, c_f64p, C.c_void_p]),
    "mpx_ipopt_eval_jac_g": (C.c_int, [C.c_int, c_f64p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), c_f64p, C.c_void_p]),
    "mpx_ipopt_eval_h": (C.c_int, [C.c_int, c_f64p, C.c_int, C.c_double, C.c_int, c_f64p, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), c_f64p, C.c_void_p]),
    "mpx_casadi_bind": (C.c_int, [C.c_void_p]),
    "mpx_sync": (C.c_int, [C.c_void_p]),
    "mpx_peer_alloc": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "mpx_peer_open": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mpx_peer_close": (C.c_int, [C.c_void_p]),
    "mpx_peer_free": (C.c_int, [C.c_void_p]),
    "mpx_eval_g_jac_dev_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "mpx_launch_count": (C.c_int64, [C.c_void_p]),
    "mpx_program_origin": (C.c_char_p, [C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). mpopt_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != MPX_OK:
        raise MpxError(rc, lib().mpx_last_error().decode())


def ptr(a, typ=c_f64p):
    return None if a is None else a.ctypes.data_as(typ)
