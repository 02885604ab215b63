"""Committed golden vectors (tests/golden/*.npz, produced from the oracle by tests/golden/make_golden.py):
the oracle must keep reproducing them on CPU; the CUDA path must reproduce them on the GPU."""
import os

import numpy as np
import pytest

from helpers import assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
sys_path_golden = os.path.join(HERE, "golden")


def _cases():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(sys_path_golden, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


_M = _cases()
CASES = _M.CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    G = np.load(os.path.join(sys_path_golden, name + ".npz"))
    ora, _, _ = _M.build(CASES[name])
    f, g, grad, J = ora._eval(G["z"], G["p"])
    assert np.array_equal(J.indptr, G["rowptr"]) and np.array_equal(J.indices, G["colind"])
    assert_close(J.data, G["values"], "values", 1e-13)
    assert_close(g, G["g"], "g", 1e-13)
    assert_close(grad, G["grad"], "grad", 1e-13)
    assert abs(f - float(G["f"])) <= 1e-13 * max(1.0, abs(float(G["f"])))
    for a, b in zip(ora.bounds(), (G["zmin"], G["zmax"], G["gmin"], G["gmax"])):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(libmpx, name):
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY

    problem, K, po, scheme = CASES[name][:4]
    G = np.load(os.path.join(sys_path_golden, name + ".npz"))
    tr = Transcription(REGISTRY[problem](), K, po, scheme, drop_exact_zeros=False, adaptive=len(CASES[name]) > 4)
    rp, ci = tr.structure()
    assert np.array_equal(rp, G["rowptr"]) and np.array_equal(ci, G["colind"])  # bit-exact index work
    g = np.empty(tr.n_g)
    assert_close(tr.jac_g_values(G["z"], G["p"], g_out=g), G["values"], "values")  # 1e-10 relative (north_star)
    assert_close(g, G["g"], "g")
    assert_close(tr.grad_f(G["z"], G["p"]), G["grad"], "grad_f")
    assert_close(tr.f(G["z"], G["p"]), float(G["f"]), "f")
    for a, b in zip(tr.bounds(), (G["zmin"], G["zmax"], G["gmin"], G["gmax"])):
        assert np.array_equal(a, b)
    assert np.allclose(tr.initial_guess(), G["z0"], rtol=0, atol=1e-15)


HESS = sorted(n for n in CASES if os.path.exists(os.path.join(sys_path_golden, n + "_hess.npz")))


@pytest.mark.parametrize("name", HESS)
def test_oracle_reproduces_golden_hessian(name):
    from oracle.hessian import hess_l

    G = np.load(os.path.join(sys_path_golden, name + "_hess.npz"))
    ora, _, _ = _M.build(CASES[name])
    H = hess_l(ora, G["z"], G["p"], float(G["lam_f"]), G["lam_g"])
    assert np.array_equal(H.indptr, G["rowptr"]) and np.array_equal(H.indices, G["colind"])
    assert_close(H.data, G["values"], "hess_l values", 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", HESS)  # base NLPs and the widths-as-variables NLP (round 2: device Hessian)
def test_cuda_reproduces_golden_hessian(libmpx, name):
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY

    problem, K, po, scheme = CASES[name][:4]
    G = np.load(os.path.join(sys_path_golden, name + "_hess.npz"))
    tr = Transcription(REGISTRY[problem](), K, po, scheme, drop_exact_zeros=False, adaptive=len(CASES[name]) > 4)
    rp, ci = tr.hess_structure()
    assert np.array_equal(rp, G["rowptr"]) and np.array_equal(ci, G["colind"])
    assert_close(tr.hess_l_values(G["z"], G["p"], float(G["lam_f"]), G["lam_g"]), G["values"], "hess_l values")
