"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8c).  Each test cites the reference test it reproduces.  CPU only."""
import numpy as np
import pytest

import oracle as O


def test_householder_kat():  # src/householder.rs:103-118
    a = np.array([1.5, 2.0, 3.0])
    ret = O.reflection_axis(a)
    assert abs(ret - (-3.90512)) < 1e-4
    np.testing.assert_allclose(a, [0.8319, 0.3078, 0.4617], atol=1e-4)
    assert abs(a @ a - 1.0) < 1e-4
    a = np.array([-3.0, 0, 0, 0])
    ret = O.reflection_axis(a)
    assert abs(ret - 3.0) < 1e-4
    np.testing.assert_allclose(a, [-1.0, 0, 0, 0], atol=1e-4)
    a = np.array([0.0, 0.0])
    assert O.reflection_axis(a) is None
    np.testing.assert_array_equal(a, [0.0, 0.0])


def test_reflect_plane_col():  # src/reflection.rs:48-63
    y = np.array([0.0, 1.0, 0.0])
    v = np.array([[1.0, 2, 3], [3, 4, 5]]).T
    O.reflect_cols(y, v)
    np.testing.assert_array_equal(v, np.array([[1.0, -2, 3], [3, -4, 5]]).T)
    O.reflect_cols(y, v)
    np.testing.assert_array_equal(v, np.array([[1.0, 2, 3], [3, 4, 5]]).T)
    v = np.array([[1.0, 2, 3], [3, 4, 5]]).T
    O.reflect_cols(y, v, bias=3.0)
    np.testing.assert_array_equal(v, np.array([[1.0, 4, 3], [3, 2, 5]]).T)


def test_reflect_plane_row():  # src/reflection.rs:65-79
    y = np.array([0.0, 1.0, 0.0])
    v = np.array([[1.0, 2, 3], [3, 4, 5]])
    O.reflect_rows(y, v)
    np.testing.assert_array_equal(v, [[1.0, -2, 3], [3, -4, 5]])
    v = np.array([[1.0, 2, 3], [3, 4, 5]])
    O.reflect_rows(y, v, bias=3.0)
    np.testing.assert_array_equal(v, [[1.0, 4, 3], [3, 2, 5]])


def test_qr_kat():  # src/qr.rs:257-276
    a = np.array([[3.2, 1.3], [4.4, 5.2], [1.3, 6.7]])
    d = O.qr(a)
    q, r = O.generate_q(a, d), O.qr_into_r(a, d)
    np.testing.assert_allclose(q, [[0.5720674, -0.4115578], [0.7865927, 0.0301901], [0.2324024, 0.9108835]], atol=1e-5)
    np.testing.assert_allclose(r, [[5.594, 6.391], [0.0, 5.725]], atol=1e-3)
    z = np.zeros((2, 2))
    d = O.qr(z)
    np.testing.assert_array_equal(O.generate_q(z, d), np.eye(2))
    np.testing.assert_array_equal(O.qr_into_r(z, d), np.zeros((2, 2)))


def _solve(a, b):  # qr.rs:124-152 composed from oracle pieces
    a = a.copy(); b = b.copy()
    d = O.qr(a)
    O.qt_mul(a, d, b)
    n = a.shape[1]
    x = b[:n]
    O.solve_triangular(a[:n, :n], x, O.UPPER, ext_diag=np.abs(d))
    return x


def test_qr_solve_and_inverse():  # src/qr.rs:279-336
    a = np.array([[1.0, 9.80], [-7.0, 3.3]])
    x = np.array([[3.2, 1.3, 4.4], [5.2, 1.3, 6.7]])
    np.testing.assert_allclose(_solve(a, a @ x), x, atol=1e-5)
    a3 = np.array([[3.2, 1.3], [4.4, 5.2], [1.3, 6.7]])
    np.testing.assert_allclose(_solve(a3, a3 @ x), x, atol=1e-5)
    np.testing.assert_allclose(_solve(a, np.eye(2)), [[0.04589, -0.1363], [0.09735, 0.0139]], atol=1e-4)


def test_qt_mul():  # src/qr.rs:363-380
    a = np.array([[1.0, 9.80], [-7.0, 3.3]])
    b = np.array([[3.2, 1.3, 4.4], [5.2, 1.3, 6.7]])
    d = O.qr(a)
    res = O.generate_q(a, d).T @ b
    O.qt_mul(a, d, b)
    np.testing.assert_allclose(b, res, atol=1e-7)


def test_inverse_scaled_identity():  # tests/qr.rs:55-75
    a = np.eye(5) * 1e-20
    inv = _solve(a, np.eye(5))
    np.testing.assert_allclose(inv, np.eye(5) * 1e20, atol=1e-3)


def test_cholesky_kat():  # src/cholesky.rs:209-245, tests/cholesky.rs:87-95
    for dt in (np.float64, np.float32):
        a = np.array([[25.0, 15, -5], [15, 18, 0], [-5, 0, 11]], dtype=dt)
        st, _ = O.cholesky(a)
        assert st == 0
        np.testing.assert_allclose(a, [[5.0, 0, 0], [3, 3, 0], [-1, 1, 3]], atol=1e-7)
    assert O.cholesky(np.array([[1.0, 2], [2, 1]]))[0] == 1
    assert O.cholesky(np.zeros((2, 2)))[0] == 1
    e = np.zeros((0, 0)); assert O.cholesky(e)[0] == 0
    one = np.ones((1, 1)); assert O.cholesky(one)[0] == 0 and one[0, 0] == 1.0


def test_cholesky_dirty_keeps_upper():  # src/cholesky.rs:17-19
    rng = np.random.default_rng(1)
    g = rng.uniform(-100, 100, (7, 7)); s = g.T @ g + np.eye(7)
    d = s.copy(); O.cholesky(d, clean=False)
    np.testing.assert_array_equal(np.triu(d, 1), np.triu(s, 1))
    c = s.copy(); O.cholesky(c, clean=True)
    np.testing.assert_array_equal(np.tril(d), c)


def test_solvec():  # src/cholesky.rs:247-259
    a = np.array([[25.0, 15, -5], [15, 18, 0], [-5, 0, 11]])
    x = np.array([[10.0, -3, 2.2, 4], [0, 2.4, -0.9, 1.1], [5.5, 7.6, 8.1, 10]])
    b = a @ x
    O.cholesky(a, clean=False)
    O.solve_triangular(a, b, O.LOWER)
    O.solve_triangular(a.T, b, O.UPPER)
    np.testing.assert_allclose(b, x, atol=1e-7)


def test_triangular_kat():  # src/triangular.rs:233-276
    sq = np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]])
    u = sq.copy(); O.triangular_inplace(u, O.UPPER)
    np.testing.assert_array_equal(u, [[1, 2, 3], [0, 5, 6], [0, 0, 9]])
    l = sq.copy(); O.triangular_inplace(l, O.LOWER)
    np.testing.assert_array_equal(l, [[1, 0, 0], [4, 5, 0], [7, 8, 9]])
    lower = np.array([[1.0, 0.0], [3.0, 4.0]])
    exp = np.array([[2.2, 3.1, 2.2], [1.0, 0.0, 5.7]])
    b = lower @ exp; O.solve_triangular(lower, b, O.LOWER)
    np.testing.assert_allclose(b, exp, atol=1e-7)
    upper = np.array([[4.4, 2.1], [0.0, 4.3]])
    b = upper @ exp; O.solve_triangular(upper, b, O.UPPER)
    np.testing.assert_allclose(b, exp, atol=1e-7)
    dz = np.array([[0.0, 3], [2, 0]]); z = np.zeros((2, 2))
    with np.errstate(all="ignore"):
        O.solve_triangular(dz, z, O.LOWER)  # zero diagonal must not crash


def test_triangular_known_failure():  # tests/triangular.rs:49-180
    from golden_vectors import TRI_KNOWN_A, TRI_KNOWN_X
    a, x = np.array(TRI_KNOWN_A), np.array(TRI_KNOWN_X)
    b = a @ x
    O.solve_triangular(a, b, O.UPPER)
    np.testing.assert_allclose(b, x, atol=1e-4)


def test_tridiagonal_kat():  # src/tridiagonal.rs:124-152
    arr = np.array([[4.0, 1, -2, 2], [1, 2, 0, 1], [-2, 0, 3, -2], [2, 1, -2, -1]])
    a = arr.copy()
    off = O.sym_tridiagonal(a)
    np.testing.assert_allclose(np.diag(a), [4, 10 / 3, -33 / 25, 149 / 75], atol=1e-5)
    np.testing.assert_allclose(np.abs(off), [3, 5 / 3, 68 / 75], atol=1e-5)
    q = O.assemble_q(a, 1, off)
    t = np.diag(np.diag(a)) + np.diag(np.abs(off), 1) + np.diag(np.abs(off), -1)
    np.testing.assert_allclose(q @ t @ q.T, arr, atol=1e-9)
    np.testing.assert_allclose(q @ q.T, np.eye(4), atol=1e-9)
    one = np.array([[1.1]])
    assert O.sym_tridiagonal(one).size == 0 and one[0, 0] == 1.1


@pytest.mark.parametrize("arr", [
    np.array([[4.0, 0, 2, 2], [-2, 6, 3, -2], [2, 7, -3.2, -1]]),          # src/bidiagonal.rs:142-167 (lower)
    np.array([[4.0, 0, 2], [-2, 6, 3], [2, 7, -3.2], [4, -3, 0.2]]),       # src/bidiagonal.rs:169-190 (upper)
])
def test_bidiagonal_kat(arr):
    a = arr.copy()
    d, e = O.bidiagonal(a)
    rows, cols = arr.shape
    upper = rows >= cols
    u = O.assemble_q(a, 0 if upper else 1, d if upper else e)
    vt = O.assemble_q(a.T, 1 if upper else 0, e if upper else d).T
    md = min(rows, cols)
    b = np.diag(np.abs(d)) + (np.diag(np.abs(e), 1) if upper else np.diag(np.abs(e), -1))
    assert u.shape == (rows, md) and vt.shape == (md, cols)
    np.testing.assert_allclose(u.T @ u if upper else u @ u.T, np.eye(md), atol=1e-5)
    np.testing.assert_allclose(vt @ vt.T, np.eye(md), atol=1e-5)
    np.testing.assert_allclose(u @ b @ vt, arr, atol=1e-5)


# ---- eigh.rs KATs (the Givens phase on top of sym_tridiagonal) -------------------------------------
def test_eigh_kats():
    # src/eigh.rs:357-372 symm_eigvals
    vals, vecs = O.symmetric_eig(np.array([[6.0, 2], [2, 6]]), vectors=False)
    np.testing.assert_allclose(vals, [8, 4], atol=1e-12)
    assert vecs is None
    vals, _ = O.symmetric_eig(np.array([[1.0, -5, 7], [-5, 2, -9], [7, -9, 3]]), vectors=False)
    np.testing.assert_allclose(np.sort(vals), [-6.86819, -3.41558, 16.28378], atol=1e-5)
    # src/eigh.rs:374-409 sym_eigvecs1..3
    for a, exp in (([[3.0, 1, 1], [1, 3, 1], [1, 1, 3]], [5, 2, 2]), ([[6.0, 2], [2, 6]], [8, 4]),
                   ([[1.0, -5, 7], [-5, 2, -9], [7, -9, 3]], [16.28378, -3.41558, -6.86819])):
        a = np.array(a)
        vals, vecs = O.symmetric_eig(a.copy(), vectors=True)
        order = np.argsort(-vals)
        np.testing.assert_allclose(vals[order], exp, atol=1e-5)
        np.testing.assert_allclose(vecs.T @ vecs, np.eye(len(exp)), atol=1e-5)
        np.testing.assert_allclose(a @ vecs, vecs * vals[None, :], atol=1e-5)
    # tests/eigh.rs:59-65 eigh_f32
    v32, _ = O.symmetric_eig(np.array([[1, -5, 7], [-5, 2, -9], [7, -9, 3]], dtype=np.float32), vectors=False)
    np.testing.assert_allclose(v32, [16.28378, -3.41558, -6.86819], atol=1e-5)
    # 1 x 1 and a random symmetric matrix against LAPACK
    vals, vecs = O.symmetric_eig(np.array([[2.5]]))
    assert vals[0] == 2.5 and vecs[0, 0] == 1.0
    g = np.random.default_rng(0).uniform(-100, 100, (40, 40))
    s = (g + g.T) / 2
    vals, vecs = O.symmetric_eig(s.copy())
    np.testing.assert_allclose(np.sort(vals), np.linalg.eigvalsh(s), atol=1e-10)
    np.testing.assert_allclose(s @ vecs, vecs * vals[None, :], atol=1e-9)


# ---- svd.rs KATs (the Givens phase on top of bidiagonal) -------------------------------------------
def _sort_svd(u, s, vt, desc=True):
    idx = np.argsort(-s if desc else s, kind="stable")
    return (None if u is None else u[:, idx]), s[idx], (None if vt is None else vt[idx, :])


def test_svd_kats():
    # src/svd.rs:537-553 svd_test
    u, s, vt = _sort_svd(*O.svd(np.array([[3.0, 0], [0, -2]]), eps=1e-15))
    np.testing.assert_allclose(s, [3, 2], atol=1e-7)
    np.testing.assert_allclose(u, [[1, 0], [0, -1]], atol=1e-7)
    np.testing.assert_allclose(vt, [[1, 0], [0, 1]], atol=1e-7)
    u, s, vt = O.svd(np.array([[1.0, 0, -1], [-2, 1, 4]]), False, False, eps=1e-15)
    np.testing.assert_allclose(s, [0.51371, 4.76824], atol=1e-5)          # unsorted, as the reference's test
    assert u is None and vt is None
    # src/svd.rs:555-600 svd_props
    big = np.array([[10.74785316637712, -5.994983325167452, -6.064492921857296],
                    [-4.149751381521569, 20.654504205822462, -4.470436210703133],
                    [-22.772715014220207, -1.4554372570788008, 18.108113992170573]]).T
    for a, exp in ((np.array([[-2.0, 1, 4]]), [np.sqrt(21)]), (np.array([[1.0, 1], [1, 1]]), [2, 0]),
                   (np.array([[-3.0, 4], [4.3, 2.1], [6.6, 8.7]]), [11.80876, 5.2633658]),
                   (big, [3.16188022e+01, 2.23811978e+01, 0])):
        u, s, vt = _sort_svd(*O.svd(a.copy(), eps=1e-15))
        np.testing.assert_allclose(s, exp, atol=1e-5)
        assert not np.any(np.signbit(s))
        np.testing.assert_allclose(u @ np.diag(s) @ vt, a, atol=1e-5)
        for cu, cv in ((False, True), (True, False), (False, False)):
            u2, s2, vt2 = _sort_svd(*O.svd(a.copy(), cu, cv, eps=1e-15))
            np.testing.assert_allclose(s2, s, atol=1e-9)
            if cu:
                np.testing.assert_allclose(u2, u, atol=1e-9)
            if cv:
                np.testing.assert_allclose(vt2, vt, atol=1e-9)
    # src/svd.rs:602-614 svd_corner, tests/svd.rs:66-72 svd_f32
    u, s, vt = O.svd(np.array([[0.0]]), eps=1e-15)
    assert s[0] == 0 and u[0, 0] == 1 and vt[0, 0] == 1
    u, s, vt = O.svd(np.array([[3, 0], [0, -2]], dtype=np.float32))
    np.testing.assert_allclose(s, [3, 2], atol=1e-7)
    np.testing.assert_allclose(u, [[1, 0], [0, -1]], atol=1e-7)
    np.testing.assert_allclose(vt, [[1, 0], [0, 1]], atol=1e-7)
    # random rectangular matrices against LAPACK + the reference's properties (tests/svd.rs:10-30)
    for shape in ((30, 30), (50, 20), (20, 50), (1, 9), (9, 1)):
        a = np.random.default_rng(shape[0] * 64 + shape[1]).uniform(-100, 100, shape)
        u, s, vt = O.svd(a.copy())
        np.testing.assert_allclose(np.sort(s)[::-1], np.linalg.svd(a, compute_uv=False), atol=1e-9)
        np.testing.assert_allclose(u @ np.diag(s) @ vt, a, atol=1e-9)
        k = min(shape)
        np.testing.assert_allclose(vt @ vt.T, np.eye(k), atol=1e-10)
        np.testing.assert_allclose(u.T @ u if shape[0] >= shape[1] else u @ u.T, np.eye(k), atol=1e-10)


def test_lobpcg_orthonormalize_properties():  # src/lobpcg/algorithm.rs:486-503 (test_orthonormalize), same tolerances
    m = np.random.default_rng(0).uniform(0, 1, (10, 10)) * 10.0
    st, fi, n, l = O.lobpcg_orthonormalize(m.copy())
    assert st == 0 and fi == -1
    np.testing.assert_allclose(n @ n.T, np.eye(10), atol=1e-2)
    w = m.copy()
    d = O.qr(w)
    np.testing.assert_allclose(np.abs(O.qr_into_r(w, d)), np.abs(l.T), atol=1e-2)
    assert np.all(np.triu(l, 1) == 0)                       # cholesky_into zeroes the strict upper (cholesky.rs:78-82)
    # a rank-deficient block fails exactly where cholesky_into does (algorithm.rs:83 `?`)
    bad = m.copy()
    bad[:, 3] = 0.0
    st, fi, _, _ = O.lobpcg_orthonormalize(bad)
    assert st == 1 and fi == 3


def test_lobpcg_apply_constraints_restatement():  # src/lobpcg/algorithm.rs:63-76
    rng = np.random.default_rng(1)
    y = np.linalg.qr(rng.uniform(-1, 1, (30, 4)))[0]        # orthonormal constraints: cholesky_yy = I and v becomes orthogonal to y
    v = rng.uniform(-1, 1, (30, 5))
    lyy = (y.T @ y).copy()
    assert O.cholesky(lyy)[0] == 0
    out = O.lobpcg_apply_constraints(v.copy(), lyy, y)
    assert np.max(np.abs(y.T @ out)) <= 1e-13
    # general y: exactly the composition the reference performs (one forward solve, no back solve)
    y2 = rng.uniform(-1, 1, (30, 4))
    l2 = (y2.T @ y2).copy()
    assert O.cholesky(l2)[0] == 0
    out2 = O.lobpcg_apply_constraints(v.copy(), l2, y2)
    np.testing.assert_allclose(out2, v - y2 @ np.linalg.solve(np.tril(l2), y2.T @ v), atol=1e-12)
