"""Oracle of the Lagrangian Hessian against central finite differences of the first-order oracle
(grad_f and jac_g of oracle/nlp.py, themselves checked by finite differences in test_oracle_nlp.py)."""
import numpy as np
import pytest

from helpers import random_point
from oracle.hessian import hess_l
from oracle.nlp import OracleNLP


@pytest.mark.parametrize("problem,K,po,scheme", [("moon_lander", 3, 3, "LGR"), ("kitchen_sink", 3, [3, 2, 4], "LGL"),
                                                 ("two_phase_schwartz", 2, 4, "LGR"), ("van_der_pol", 3, [2, 5, 3], "CGL"),
                                                 ("robot_arm", 2, 3, "LGR"), ("synthetic_6_3", 2, 4, "LGR"),
                                                 ("delta3_launch_vehicle", 1, 3, "LGR")])
def test_hessian_by_finite_differences(problem, K, po, scheme):
    from mpopt_b200.problems import REGISTRY

    ora = OracleNLP(REGISTRY[problem](), K, po, scheme)
    z, p = random_point(ora, dirichlet=True)
    if problem == "delta3_launch_vehicle":
        z = ora.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(7).standard_normal(ora.n_z))
    rng = np.random.default_rng(9)
    lam, sig = rng.uniform(-1, 1, ora.n_g), 0.7
    H = hess_l(ora, z, p, sig, lam)
    assert (H.tocoo().row >= H.tocoo().col).all()
    Hs = (H + sp_tril_strict_T(H)).toarray()

    def grad_lag(zz):
        return sig * ora.grad_f(zz, p) + ora.jac_g(zz, p).T @ lam

    e = 1e-6
    for j in rng.choice(ora.n_z, size=min(ora.n_z, 25), replace=False):
        dz = np.zeros(ora.n_z)
        dz[j] = e
        col = (grad_lag(z + dz) - grad_lag(z - dz)) / (2 * e)
        assert np.abs(col - Hs[:, j]).max() <= 1e-5 * max(1.0, np.abs(col).max()), (j, np.abs(col - Hs[:, j]).max())
    # structural entries only where something can be non-zero: the FD Hessian has no entry outside the pattern
    pat = (Hs != 0) | (H + sp_tril_strict_T(H)).astype(bool).toarray()
    for j in rng.choice(ora.n_z, size=min(ora.n_z, 10), replace=False):
        dz = np.zeros(ora.n_z)
        dz[j] = e
        col = (grad_lag(z + dz) - grad_lag(z - dz)) / (2 * e)
        assert np.abs(col[~pat[:, j]]).max(initial=0.0) <= 1e-6


def sp_tril_strict_T(H):
    import scipy.sparse as sp

    return sp.tril(H, k=-1).T.tocsr()
