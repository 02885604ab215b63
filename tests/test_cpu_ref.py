"""The compiled CPU baseline (oracle/cpu_ref, test infrastructure) against the numpy oracle: same CSR structure,
values and constraint vector to 1e-13 -- so that the number bench.py reports as ``cpu_baseline`` / ``--impl reference``
is the time of a CORRECT evaluation of the same NLP."""
import numpy as np
import pytest

from mpopt_b200.problems import REGISTRY
from oracle.cpu_ref import CpuRef, synthetic_params
from oracle.nlp import OracleNLP

CASES = [
    ("moon_lander", 20, 3, "LGR"),            # BASELINE.json configs[0]
    ("moon_lander", 7, 15, "LGR"),
    ("synthetic_6_3", 9, 15, "LGR"),          # headline shape
    ("synthetic_6_3", 5, 20, "LGL"),          # configs[3] shape
    ("van_der_pol", 12, [30 if k % 3 == 1 else 3 for k in range(12)], "CGL"),  # configs[2] shape
    ("van_der_pol", 3, 4, "LGL"),
]


@pytest.mark.parametrize("problem,K,po,scheme", CASES)
def test_cpu_ref_matches_oracle(problem, K, po, scheme):
    ocp = REGISTRY[problem]()
    ora = OracleNLP(ocp, K, po, scheme, drop_exact_zeros=False)
    ref = CpuRef(problem, K, po, scheme, midu=True, params=synthetic_params() if problem == "synthetic_6_3" else None)
    assert (ref.n_z, ref.n_g, ref.nnz) == (ora.n_z, ora.n_g, ora.jac_g(np.zeros(ora.n_z) + 0.1).nnz)
    rng = np.random.default_rng(3)
    z = rng.uniform(-1, 1, ora.n_z)
    z[-2:] = [0.1, 2.3]
    w = rng.dirichlet(np.ones(K))
    J = ora.jac_g(z, w)
    rp, ci = ref.structure()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
    g, vals = ref.eval(z, w)
    assert np.max(np.abs(g - ora.g(z, w)) / np.maximum(1, np.abs(g))) < 1e-13
    assert np.max(np.abs(vals - J.data) / np.maximum(1, np.abs(J.data))) < 1e-13


def test_cpu_ref_thread_count_does_not_change_results():
    ref = CpuRef("synthetic_6_3", 64, 15, "LGR", params=synthetic_params())
    rng = np.random.default_rng(0)
    z = rng.uniform(-1, 1, ref.n_z)
    z[-2:] = [0.0, 1.0]
    w = np.full(64, 1 / 64)
    ref.set_threads(1)
    g1, v1 = ref.eval(z, w)
    ref.set_threads(max(2, ref.threads))
    g2, v2 = ref.eval(z, w)
    assert np.array_equal(g1, g2) and np.array_equal(v1, v2)
