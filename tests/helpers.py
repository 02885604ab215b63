"""Shared helpers for the parity tests."""
import numpy as np

RTOL = 1e-10  # north_star: float64 residual / Jacobian values within 1e-10 relative


def assert_close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert not np.isnan(a).any(), f"{what}: NaN in result"
    assert err.size == 0 or err.max() <= rtol, f"{what}: max rel err {err.max():.3e} at {int(err.argmax())}"


def random_point(oracle, seed=20261017, tf=None, dirichlet=False):
    """Seeded evaluation point in the layout of z, plus segment widths (SURVEY.md 8d 'Input values')."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, oracle.n_z)
    for ph in range(oracle.P):
        z[oracle.colT0(ph)] = 0.25 * ph + (0.1 if ph else 0.0)
        z[oracle.colTF(ph)] = (tf if tf is not None else 2.0) + 1.5 * ph
        for m in range(oracle.na):
            z[oracle.colA(ph, m)] = rng.uniform(0.2, 1.2)
    if dirichlet:
        p = np.concatenate([rng.dirichlet(np.ones(oracle.K)) for _ in range(oracle.P)])
    else:
        p = oracle.seg_width_params()
    return z, p
