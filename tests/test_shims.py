"""Solver-facing shims of include/mpx.h: IPOPT's C-interface callbacks and CasADi's external-function ABI.

CPU part: every symbol is exported and refuses to compute without a bound plan.  GPU part: the callbacks, driven
through their C calling conventions exactly as IPOPT / CasADi would, return what the oracle computes (the reference's
nlp_f / nlp_grad_f / nlp_g / nlp_jac_g, /root/reference/mpopt/mpopt.py:757, :804)."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_close, random_point


def test_casadi_symbols_exported(libmpx):
    from mpopt_b200 import _lib

    for name in _lib.CASADI_FUNCTIONS:
        for suf in _lib.CASADI_SUFFIXES:
            assert getattr(libmpx, name + suf) is not None
    libmpx.nlp_jac_g_n_in.restype = libmpx.nlp_jac_g_n_out.restype = C.c_longlong
    assert libmpx.nlp_jac_g_n_in() == 2 and libmpx.nlp_jac_g_n_out() == 2
    libmpx.nlp_grad_f_name_out.restype = C.c_char_p
    libmpx.nlp_grad_f_name_out.argtypes = [C.c_longlong]
    assert libmpx.nlp_grad_f_name_out(1) == b"grad_f_x"
    # nothing bound: no sparsity, evaluation reports failure (non-zero), nothing is computed on the host
    assert libmpx.mpx_casadi_bind(None) == 0
    libmpx.nlp_g_sparsity_in.restype = C.c_void_p
    libmpx.nlp_g_sparsity_in.argtypes = [C.c_longlong]
    assert libmpx.nlp_g_sparsity_in(0) is None
    assert libmpx.nlp_g(None, None, None, None, 0) != 0


def test_ipopt_callbacks_reject_missing_plan(libmpx):
    from mpopt_b200 import _lib

    x, out = np.zeros(4), np.zeros(4)
    assert libmpx.mpx_ipopt_eval_f(4, _lib.ptr(x), 1, _lib.ptr(out), None) == 0
    d = _lib.IpoptData(None, None)
    assert libmpx.mpx_ipopt_eval_g(4, _lib.ptr(x), 1, 4, _lib.ptr(out), C.byref(d)) == 0
    assert libmpx.mpx_stage(None, None, None, 0) == _lib.MPX_EINVAL
    assert libmpx.mpx_fetch(None, 4, None) == _lib.MPX_EINVAL


def _setup(make, K, p, scheme):
    from mpopt_b200.nlp import Transcription
    from oracle.nlp import OracleNLP

    ocp = make()
    tr = Transcription(ocp, K, p, scheme, drop_exact_zeros=False)
    ora = OracleNLP(ocp, K, p, scheme, drop_exact_zeros=False)
    z, w = random_point(ora, dirichlet=True)
    return tr, ora, z, w


@pytest.mark.gpu
def test_ipopt_callbacks_match_oracle(libmpx):
    from mpopt_b200 import _lib
    from mpopt_b200.problems import kitchen_sink

    tr, ora, z, w = _setup(kitchen_sink, 4, [3, 5, 4, 3], "LGR")
    n, m, nnz = tr.n_z, tr.n_g, tr.nnz
    d = _lib.IpoptData(tr._plan, _lib.ptr(w))
    ud = C.byref(d)
    # structure request (values == NULL): triplets in CSR order, C indexing
    ir, jc = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    assert libmpx.mpx_ipopt_eval_jac_g(n, _lib.ptr(z), 1, m, nnz, ip(ir), ip(jc), None, ud) == 1
    J = ora.jac_g(z, w)
    assert np.array_equal(jc, J.indices) and np.array_equal(ir, np.repeat(np.arange(m), np.diff(J.indptr)))
    # IPOPT's call order for one iterate: f (new_x), grad_f, g, jac_g (same x) -> one fused evaluation
    f, grad, g, vals = np.zeros(1), np.zeros(n), np.zeros(m), np.zeros(nnz)
    l0 = tr.launches
    assert libmpx.mpx_ipopt_eval_f(n, _lib.ptr(z), 1, _lib.ptr(f), ud) == 1
    l1 = tr.launches
    assert libmpx.mpx_ipopt_eval_grad_f(n, _lib.ptr(z), 0, _lib.ptr(grad), ud) == 1
    assert libmpx.mpx_ipopt_eval_g(n, _lib.ptr(z), 0, m, _lib.ptr(g), ud) == 1
    assert libmpx.mpx_ipopt_eval_jac_g(n, _lib.ptr(z), 0, m, nnz, None, None, _lib.ptr(vals), ud) == 1
    assert l1 > l0 and tr.launches == l1, "new_x = 0 must not launch anything"
    assert abs(f[0] - ora.f(z, w)) <= 1e-10 * max(1.0, abs(ora.f(z, w)))
    assert_close(grad, ora.grad_f(z, w), "grad_f")
    assert_close(g, ora.g(z, w), "g")
    assert_close(vals, J.data, "jac_g values")
    # Eval_H_CB: pattern request, then values
    from oracle.hessian import hess_l

    lam = np.random.default_rng(1).uniform(-1, 1, m)
    H = hess_l(ora, z, w, 0.8, lam)
    nh = H.nnz
    hr, hc, hv = np.zeros(nh, np.int32), np.zeros(nh, np.int32), np.zeros(nh)
    assert libmpx.mpx_ipopt_eval_h(n, _lib.ptr(z), 0, 0.8, m, _lib.ptr(lam), 1, nh, ip(hr), ip(hc), None, ud) == 1
    assert np.array_equal(hc, H.indices) and np.array_equal(hr, np.repeat(np.arange(n), np.diff(H.indptr)))
    assert libmpx.mpx_ipopt_eval_h(n, _lib.ptr(z), 0, 0.8, m, _lib.ptr(lam), 1, nh, None, None, _lib.ptr(hv), ud) == 1
    assert_close(hv, H.data, "eval_h values")
    # a new point
    z2 = z + 1e-3
    assert libmpx.mpx_ipopt_eval_g(n, _lib.ptr(z2), 1, m, _lib.ptr(g), ud) == 1
    assert_close(g, ora.g(z2, w), "g at the second point")
    # wrong sizes are refused
    assert libmpx.mpx_ipopt_eval_g(n + 1, _lib.ptr(z), 1, m, _lib.ptr(g), ud) == 0


@pytest.mark.gpu
def test_ipopt_eval_h_first_at_a_new_iterate_restages(libmpx):
    """eval_h may be the first callback at a new x: the callbacks that follow with new_x = 0 must see THAT x."""
    from mpopt_b200 import _lib
    from mpopt_b200.problems import kitchen_sink
    from oracle.hessian import hess_l

    tr, ora, z, w = _setup(kitchen_sink, 4, [3, 5, 4, 3], "LGR")
    n, m, nnz = tr.n_z, tr.n_g, tr.nnz
    d = _lib.IpoptData(tr._plan, _lib.ptr(w))
    ud = C.byref(d)
    f, g, vals = np.zeros(1), np.zeros(m), np.zeros(nnz)
    assert libmpx.mpx_ipopt_eval_f(n, _lib.ptr(z), 1, _lib.ptr(f), ud) == 1  # iterate 1
    z2 = z + 0.01 * np.random.default_rng(2).standard_normal(n)
    lam = np.random.default_rng(1).uniform(-1, 1, m)
    H = hess_l(ora, z2, w, 0.7, lam)
    hv = np.zeros(H.nnz)
    assert libmpx.mpx_ipopt_eval_h(n, _lib.ptr(z2), 1, 0.7, m, _lib.ptr(lam), 1, H.nnz, None, None, _lib.ptr(hv), ud) == 1
    assert_close(hv, H.data, "eval_h at the new iterate")
    assert libmpx.mpx_ipopt_eval_g(n, _lib.ptr(z2), 0, m, _lib.ptr(g), ud) == 1
    assert libmpx.mpx_ipopt_eval_jac_g(n, _lib.ptr(z2), 0, m, nnz, None, None, _lib.ptr(vals), ud) == 1
    assert libmpx.mpx_ipopt_eval_f(n, _lib.ptr(z2), 0, _lib.ptr(f), ud) == 1
    assert_close(g, ora.g(z2, w), "g after eval_h(new_x)")
    assert_close(vals, ora.jac_g(z2, w).data, "jac_g after eval_h(new_x)")
    assert abs(f[0] - ora.f(z2, w)) <= 1e-10 * max(1.0, abs(ora.f(z2, w)))


@pytest.mark.gpu
def test_host_evaluation_invalidates_staged_results(libmpx):
    """stage at z, a direct mpx_eval_g at another z, then a fetch: the stale staged result must not come back."""
    from mpopt_b200 import _lib
    from mpopt_b200.problems import kitchen_sink

    tr, ora, z, w = _setup(kitchen_sink, 4, [3, 5, 4, 3], "LGR")
    n, m = tr.n_z, tr.n_g
    assert libmpx.mpx_stage(tr._plan, _lib.ptr(z), _lib.ptr(w), _lib.MPX_STAGE_G | _lib.MPX_STAGE_JAC) == 0
    assert libmpx.mpx_staged(tr._plan) & _lib.MPX_STAGE_G
    z2 = z + 0.05
    g2 = tr.g(z2, w)
    assert_close(g2, ora.g(z2, w), "direct g")
    assert libmpx.mpx_staged(tr._plan) == 0
    g = np.zeros(m)
    assert libmpx.mpx_fetch(tr._plan, _lib.MPX_STAGE_G, _lib.ptr(g)) == _lib.MPX_EINVAL
    # the IPOPT callback with new_x = 0 notices and re-stages with the x it is handed
    d = _lib.IpoptData(tr._plan, _lib.ptr(w))
    assert libmpx.mpx_ipopt_eval_g(n, _lib.ptr(z), 0, m, _lib.ptr(g), C.byref(d)) == 1
    assert_close(g, ora.g(z, w), "g re-staged")


@pytest.mark.gpu
def test_casadi_externals_match_oracle(libmpx):
    from mpopt_b200 import _lib
    from mpopt_b200.problems import two_phase_schwartz

    tr, ora, z, w = _setup(two_phase_schwartz, 3, 4, "LGL")
    n, m, nnz = tr.n_z, tr.n_g, tr.nnz
    assert libmpx.mpx_casadi_bind(tr._plan) == 0
    try:
        # sparsity of jac_g_x: compact CCS [nrow, ncol, colind, row] == the oracle's pattern
        libmpx.nlp_jac_g_sparsity_out.restype = C.POINTER(C.c_longlong)
        libmpx.nlp_jac_g_sparsity_out.argtypes = [C.c_longlong]
        sp = libmpx.nlp_jac_g_sparsity_out(1)
        Jc = ora.jac_g(z, w).tocsc()
        Jc.sort_indices()
        assert (sp[0], sp[1]) == (m, n)
        colind = np.array([sp[2 + i] for i in range(n + 1)])
        rows = np.array([sp[2 + n + 1 + i] for i in range(nnz)])
        assert np.array_equal(colind, Jc.indptr) and np.array_equal(rows, Jc.indices)
        sx = libmpx.nlp_jac_g_sparsity_out(0)
        assert (sx[0], sx[1], sx[3]) == (m, 1, m)
        # evaluation through (arg, res, iw, w, mem)
        arg = (C.POINTER(C.c_double) * 2)(_lib.ptr(z), _lib.ptr(w))
        g, jv, f, grad = np.zeros(m), np.zeros(nnz), np.zeros(1), np.zeros(n)
        res = (C.POINTER(C.c_double) * 2)(_lib.ptr(g), _lib.ptr(jv))
        assert libmpx.nlp_jac_g(arg, res, None, None, 0) == 0
        assert_close(g, ora.g(z, w), "nlp_jac_g: g")
        assert_close(jv, Jc.data, "nlp_jac_g: jac_g_x (CCS order)")
        res = (C.POINTER(C.c_double) * 2)(_lib.ptr(f), _lib.ptr(grad))
        assert libmpx.nlp_grad_f(arg, res, None, None, 0) == 0
        assert abs(f[0] - ora.f(z, w)) <= 1e-10 * max(1.0, abs(ora.f(z, w)))
        assert_close(grad, ora.grad_f(z, w), "nlp_grad_f")
        g[:] = 0
        res1 = (C.POINTER(C.c_double) * 1)(_lib.ptr(g))
        assert libmpx.nlp_g(arg, res1, None, None, 0) == 0
        assert_close(g, ora.g(z, w), "nlp_g")
        f[:] = 0
        res1 = (C.POINTER(C.c_double) * 1)(_lib.ptr(f))
        assert libmpx.nlp_f(arg, res1, None, None, 0) == 0
        assert abs(f[0] - ora.f(z, w)) <= 1e-10 * max(1.0, abs(ora.f(z, w)))
        # nlp_hess_l (x, p, lam_f, lam_g) -> triu in CCS == the oracle's tril in CSR
        from oracle.hessian import hess_l

        lam, lf = np.random.default_rng(2).uniform(-1, 1, m), np.array([0.9])
        H = hess_l(ora, z, w, 0.9, lam)
        libmpx.nlp_hess_l_sparsity_out.restype = C.POINTER(C.c_longlong)
        libmpx.nlp_hess_l_sparsity_out.argtypes = [C.c_longlong]
        sph = libmpx.nlp_hess_l_sparsity_out(0)
        assert (sph[0], sph[1]) == (n, n)
        assert np.array_equal(np.array([sph[2 + i] for i in range(n + 1)]), H.indptr)
        assert np.array_equal(np.array([sph[3 + n + i] for i in range(H.nnz)]), H.indices)
        hv = np.zeros(H.nnz)
        arg4 = (C.POINTER(C.c_double) * 4)(_lib.ptr(z), _lib.ptr(w), _lib.ptr(lf), _lib.ptr(lam))
        res1 = (C.POINTER(C.c_double) * 1)(_lib.ptr(hv))
        assert libmpx.nlp_hess_l(arg4, res1, None, None, 0) == 0
        assert_close(hv, H.data, "nlp_hess_l")
    finally:
        libmpx.mpx_casadi_bind(None)


@pytest.mark.gpu
def test_shims_with_the_adaptive_nlp(libmpx):
    """mpopt_adaptive's NLP has no parameters (mpopt.py:3190-3191): both solver interfaces accept p = NULL for such a
    plan, report n_p = 0 and return the widths-as-variables g / jac_g / grad_f; the Hessian callback reports failure."""
    from mpopt_b200 import _lib
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import van_der_pol
    from oracle.adaptive import OracleAdaptiveNLP

    po = [3, 5, 2, 4]
    tr = Transcription(van_der_pol(), 4, po, "LGR", adaptive=True, drop_exact_zeros=False)
    ora = OracleAdaptiveNLP(van_der_pol(), 4, po, "LGR", drop_exact_zeros=False)
    rng = np.random.default_rng(11)
    z = rng.uniform(-1, 1, ora.n_z)
    z[ora.colT0(0)], z[ora.colTF(0)] = 0.0, 5.0
    z[ora.colW(0, np.arange(4))] = rng.dirichlet(np.ones(4))
    n, m, nnz = tr.n_z, tr.n_g, tr.nnz
    J = ora.jac_g(z)
    # IPOPT C interface, user_data.p = NULL
    d = _lib.IpoptData(tr._plan, None)
    ud = C.byref(d)
    f, grad, g, vals = np.zeros(1), np.zeros(n), np.zeros(m), np.zeros(nnz)
    assert libmpx.mpx_ipopt_eval_f(n, _lib.ptr(z), 1, _lib.ptr(f), ud) == 1
    assert libmpx.mpx_ipopt_eval_grad_f(n, _lib.ptr(z), 0, _lib.ptr(grad), ud) == 1
    assert libmpx.mpx_ipopt_eval_g(n, _lib.ptr(z), 0, m, _lib.ptr(g), ud) == 1
    assert libmpx.mpx_ipopt_eval_jac_g(n, _lib.ptr(z), 0, m, nnz, None, None, _lib.ptr(vals), ud) == 1
    assert abs(f[0] - ora.f(z)) <= 1e-10 * max(1.0, abs(ora.f(z)))
    assert_close(grad, ora.grad_f(z), "grad_f")
    assert_close(g, ora.g(z), "g")
    assert_close(vals, J.data, "jac_g values")
    lam = np.zeros(m)
    assert libmpx.mpx_ipopt_eval_h(n, _lib.ptr(z), 0, 1.0, m, _lib.ptr(lam), 1, 1, None, None, _lib.ptr(vals), ud) == 0
    # CasADi external functions, arg[1] = NULL, p declared with zero rows
    assert libmpx.mpx_casadi_bind(tr._plan) == 0
    try:
        libmpx.nlp_g_sparsity_in.restype = C.POINTER(C.c_longlong)
        libmpx.nlp_g_sparsity_in.argtypes = [C.c_longlong]
        sp_p = libmpx.nlp_g_sparsity_in(1)
        assert sp_p[0] == 0
        arg = (C.POINTER(C.c_double) * 2)(_lib.ptr(z), None)
        g[:] = 0
        jv = np.zeros(nnz)
        res = (C.POINTER(C.c_double) * 2)(_lib.ptr(g), _lib.ptr(jv))
        assert libmpx.nlp_jac_g(arg, res, None, None, 0) == 0
        Jc = J.tocsc()
        Jc.sort_indices()
        assert_close(g, ora.g(z), "nlp_jac_g: g")
        assert_close(jv, Jc.data, "nlp_jac_g: jac_g_x (CCS order)")
    finally:
        libmpx.mpx_casadi_bind(None)
