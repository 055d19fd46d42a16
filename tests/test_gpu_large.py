"""GPU tests of the large-matrix host paths (triangular PCIe transfers, look-ahead, TSQR tree)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16


@pytest.mark.parametrize("order", ["C", "F"])
def test_cholesky_dirty_large_triangular_transfer(order):
    """n >= 2048 takes the lower-trapezoid upload/download path; the strict upper triangle must stay untouched
    (cholesky.rs:17-19) and L L^T = A."""
    import linfa_linalg_b200 as L
    n = 2500
    rng = np.random.default_rng(3)
    g = rng.uniform(-1, 1, (n, n))
    a0 = (g + g.T) / 2 + n * np.eye(n)
    a0[np.triu_indices(n, 1)] = rng.uniform(5, 6, n * (n - 1) // 2)      # garbage in the (never read) upper triangle
    a = np.array(a0, order=order)
    L.cholesky_inplace_dirty(a)
    np.testing.assert_array_equal(np.triu(a, 1), np.triu(a0, 1))
    l = np.tril(a)
    sym = np.tril(a0) + np.tril(a0, -1).T
    assert np.linalg.norm(l @ l.T - sym) <= 8 * n * EPS * np.linalg.norm(sym)
    assert np.max(np.abs(l - np.linalg.cholesky(sym))) <= 8 * n * EPS * np.linalg.norm(sym, 2)


def test_cholesky_large_not_positive_definite_index():
    import linfa_linalg_b200 as L
    n = 2300
    g = np.random.default_rng(4).uniform(-1, 1, (n, n))
    a = (g + g.T) / 2 + n * np.eye(n)
    a[1777, 1777] = -3.0
    with pytest.raises(L.NotPositiveDefinite) as ei:
        L.cholesky_inplace(a)
    assert ei.value.index == 1777


def test_qr_lookahead_matches_no_lookahead():
    """The side-stream panel factorisation must not change results (same kernels, same order per column)."""
    import linfa_linalg_b200 as L
    m, n = 1500, 1100
    a0 = np.random.default_rng(5).uniform(-1, 1, (m, n))
    e1 = L.Engine(0)
    e2 = L.Engine(0)
    e2.set_option("lookahead", 0)
    a1, a2 = a0.copy(), a0.copy()
    d1 = L.qr_into(a1, e1).diag
    d2 = L.qr_into(a2, e2).diag
    assert np.max(np.abs(a1 - a2)) <= 64 * m * EPS * np.linalg.norm(a0)
    assert np.max(np.abs(d1 - d2)) <= 64 * m * EPS * np.linalg.norm(a0)
    e1.close(); e2.close()
