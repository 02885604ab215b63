"""Pin the oracle's collocation tables against everything the reference's own tests pin
(/root/reference/tests/test_mpopt.py:333-346, :612-634, :927-1086) plus mpmath 50-digit tables."""
import mpmath as mpm
import numpy as np
import pytest

from oracle import collocation as oc

SCHEMES = ["LGR", "LGL", "CGL"]


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("tmin", [-1.0, 0.0])
def test_degree_one_known_answers(scheme, tmin):
    """tests/test_mpopt.py:927-1086: nodes = end points, l_j(tau_i) = delta_ij, D = [[-1/h, 1/h]]*2, D2 = 0."""
    r = oc.roots(scheme, 1, tmin, 1.0)
    assert (np.abs(r - np.array([tmin, 1.0])) < 1e-6).all()
    h = r[-1] - r[0]
    C = oc.interpolation_matrix(r, r)
    assert C[0, 0] == 1 and C[1, 0] == 0 and C[0, 1] == 0 and C[1, 1] == 1
    D = oc.diff_matrix(r)
    assert (np.abs(D - np.array([[-1 / h, 1 / h], [-1 / h, 1 / h]])) < 1e-6).all()
    assert (np.abs(oc.diff_matrix(r, order=2)) < 1e-6).all()


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tau_end_points_and_shapes(scheme):
    """tests/test_mpopt.py:627-634 and :333-346."""
    po = [3, 5, 3]
    t = oc.Tables(po, scheme)
    for p in set(po):
        assert len(t.roots[p]) == p + 1
        assert t.roots[p][0] == t.tau0 and t.roots[p][-1] == t.tau1
    N = sum(po) + 1
    assert t.composite_D().shape == (N, N)
    assert t.composite_W().shape == (N,)
    assert t.composite_mid_interpolation().shape == (N - 1, N)
    assert t.composite_slope_continuity().shape == (len(po) - 1, N)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("p", [2, 3, 4, 7, 15, 20, 30])
def test_invariants(scheme, p):
    """rows of D sum to 0, sum w = tau1 - tau0, LGR w0 = 0, LGL w0 = 2/(p(p+1))."""
    t = oc.Tables([p], scheme)
    assert np.abs(t.D[p].sum(axis=1)).max() < 1e-10
    assert abs(t.w[p].sum() - 2.0) < 1e-13
    assert np.abs(t.Cmid[p].sum(axis=1) - 1.0).max() < 1e-12
    if scheme == "LGR":
        assert abs(t.w[p][0]) < 1e-14
    if scheme == "LGL":
        assert abs(t.w[p][0] - 2.0 / (p * (p + 1))) < 1e-14
    # D differentiates polynomials of degree <= p exactly
    r = t.roots[p]
    for k in range(p + 1):
        assert np.abs(t.D[p] @ r**k - k * r ** max(k - 1, 0)).max() < 1e-9 * max(1, k) ** 2


def test_numerical_vs_symbolic_mode_agree():
    """tests/test_mpopt.py:612-624: np.poly1d ('numerical') tables agree with the product form to 1e-5 at LGR p=3."""
    r = oc.roots("LGR", 3)
    D = oc.diff_matrix(r)
    w = oc.quadrature_weights(r, -1.0, 1.0)
    Dn, wn = np.zeros((4, 4)), np.zeros(4)
    for j in range(4):
        pj = np.poly1d([1.0])
        for i in range(4):
            if i != j:
                pj *= np.poly1d([1, -r[i]]) / (r[j] - r[i])
        Dn[:, j] = np.polyder(pj)(r)
        pint = np.polyint(pj)
        wn[j] = pint(1.0) - pint(-1.0)
    assert np.abs(D - Dn).max() < 1e-5 and np.abs(w - wn).max() < 1e-5


@pytest.mark.parametrize("scheme,p", [("LGR", 3), ("LGR", 15), ("LGL", 8), ("LGL", 20), ("CGL", 30)])
def test_tables_against_mpmath(scheme, p):
    """50-digit adjudication of roots (as a root of the defining polynomial), D, w and Cmid."""
    mpm.mp.dps = 50
    r = oc.roots(scheme, p)

    def legendre(n, x):  # (P_n, P_n') by the three-term recurrence in 50-digit arithmetic
        p0, p1 = mpm.mpf(1), x
        if n == 0:
            return p0, mpm.mpf(0)
        for k in range(1, n):
            p0, p1 = p1, ((2 * k + 1) * x * p1 - k * p0) / (k + 1)
        return p1, n * (x * p1 - p0) / (x * x - 1)

    # "LGR" here (Jacobi alpha=1, beta=0): the roots of P^(1,0)_{p-1} together with +1 are the zeros of P_p - P_{p-1};
    # LGL interior nodes are the zeros of P_p'
    if scheme == "LGR":
        fn = lambda y: legendre(p, y)[0] - legendre(p - 1, y)[0]
    else:
        fn = lambda y: legendre(p, y)[1]
    if scheme in ("LGR", "LGL"):
        for x in r[1:-1]:
            x = mpm.mpf(float(x))
            assert abs(fn(x) / mpm.diff(fn, x)) < mpm.mpf(4e-16)  # distance to the true root
    else:
        for j, x in enumerate(r):
            assert abs(mpm.mpf(float(x)) - mpm.cos(mpm.pi * (p - j) / p)) < mpm.mpf(5e-16)
    R = [mpm.mpf(float(x)) for x in r]

    def ell(j, t):
        v = mpm.mpf(1)
        for i in range(p + 1):
            if i != j:
                v *= (t - R[i]) / (R[j] - R[i])
        return v

    t = oc.Tables([p], scheme)
    for j in (0, 1, p // 2, p):
        wj = mpm.quad(lambda y: ell(j, y), [-1, 1])
        assert abs(wj - t.w[p][j]) < 1e-14
        for i in (0, p // 3, p):
            dij = mpm.diff(lambda y: ell(j, y), R[i])
            assert abs(dij - t.D[p][i, j]) < 1e-10 * max(1, abs(dij))
        m = p // 2
        cm = ell(j, (R[m] + R[m + 1]) / 2)
        assert abs(cm - t.Cmid[p][m, j]) < 1e-13


def test_composite_weights_drop_w0_quirk():
    """Q1: compW = [w0[0], w0[1:], w1[1:], ...] (mpopt.py:4060-4062) -- sum is short by w[0] per later segment."""
    t = oc.Tables([3, 3, 3], "LGL")
    W = t.composite_W()
    assert abs(W.sum() - (3 * 2.0 - 2 * t.w[3][0])) < 1e-13


def test_composite_D_staircase():
    """mpopt.py:4030-4039: block k>=1 contributes rows 1.. of D[p_k]; the shared node's row is the earlier one."""
    t = oc.Tables([2, 3], "LGR")
    A = t.composite_D().toarray()
    assert np.array_equal(A[:3, :3], t.D[2])
    assert np.array_equal(A[3:6, 2:6], t.D[3][1:, :])
    assert np.count_nonzero(A[:3, 3:]) == 0 and np.count_nonzero(A[3:, :2]) == 0
