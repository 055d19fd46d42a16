"""GPU parity tests: the CUDA path (through the C ABI, host views of arbitrary strides) against the
CPU oracle on the same seeded inputs, plus the reference's KATs and error behaviour.

Tolerances: elementwise |X - X_oracle| <= c * n * eps * ||A||_F, backward error and
orthogonality <= c * n * eps; exact: diag(R) >= 0, triangular zeros, off >= 0, error variants.
"""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

EPS = {np.float64: 2.220446049250313e-16, np.float32: 1.1920929e-07}


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()  # fails loudly when the CUDA library / device is missing
    return L


def tol(a, c=8.0, n=None):
    dt = a.dtype.type
    n = n or max(a.shape)
    return c * n * EPS[dt] * max(np.linalg.norm(a), 1e-300)


def layouts(a, square_t=True):
    """The layouts tests/common.rs:12-43 randomises over, plus strided slices."""
    out = [("c", np.array(a, order="C")), ("f", np.array(a, order="F"))]
    out.append(("revrows", np.array(a[::-1], order="C")[::-1]))
    out.append(("revcols", np.array(a[:, ::-1], order="C")[:, ::-1]))
    big = np.zeros((2 * a.shape[0], 3 * a.shape[1]), dtype=a.dtype)
    v = big[::2, ::3]
    v[...] = a
    out.append(("strided", v))
    if square_t and a.shape[0] == a.shape[1]:
        out.append(("t", np.array(a.T, order="C").T))
    return out


def rnd(shape, dt=np.float64, seed=0, lo=-100.0, hi=100.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(dt)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (37, 19, 53), (128, 128, 128), (130, 257, 65), (256, 64, 4100), (32, 32, 20000)])
def test_gemm_f64(L, ta, tb, m, n, k):
    import ctypes as C
    import torch
    e = L.engine()
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    A = torch.rand((k, m) if ta else (m, k), dtype=torch.float64, device="cuda", generator=g) - 0.5
    B = torch.rand((n, k) if tb else (k, n), dtype=torch.float64, device="cuda", generator=g) - 0.5
    Cm = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g)
    # torch tensors are row-major: a row-major X (r x c) is a column-major X^T (c x r, ld = c).
    # Column-major C (m x n) = op(A) op(B)  <=>  we hand the library transposed storage.
    Acm = A.t().contiguous()  # storage of column-major A
    Bcm = B.t().contiguous()
    Ccm = Cm.t().contiguous()
    ref = 1.5 * ((A.t() if ta else A) @ (B.t() if tb else B)) + 0.5 * Cm
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    st = e.call("lfb_gemm_dev_f64", ta, tb, m, n, k, 1.5, C.c_void_p(Acm.data_ptr()), A.shape[0],
                C.c_void_p(Bcm.data_ptr()), B.shape[0], 0.5, C.c_void_p(Ccm.data_ptr()), m)
    e._check(st)
    torch.cuda.synchronize()
    e.set_stream(None)
    got = Ccm.t()
    assert torch.allclose(got, ref, rtol=0, atol=1e-11 * k)


def test_gemm_f32(L):
    import ctypes as C
    import torch
    e = L.engine()
    for (ta, tb, m, n, k) in [(0, 0, 70, 33, 129), (1, 0, 64, 64, 3000), (0, 1, 5, 200, 17), (1, 1, 100, 100, 100)]:
        A = torch.rand((k, m) if ta else (m, k), dtype=torch.float32, device="cuda") - 0.5
        B = torch.rand((n, k) if tb else (k, n), dtype=torch.float32, device="cuda") - 0.5
        Ccm = torch.zeros((n, m), dtype=torch.float32, device="cuda")
        ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
        torch.cuda.synchronize()
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        Acm, Bcm = A.t().contiguous(), B.t().contiguous()   # keep alive: column-major storage of A, B
        st = e.call("lfb_gemm_dev_f32", ta, tb, m, n, k, 1.0, C.c_void_p(Acm.data_ptr()), A.shape[0],
                    C.c_void_p(Bcm.data_ptr()), B.shape[0], 0.0, C.c_void_p(Ccm.data_ptr()), m)
        e._check(st)
        torch.cuda.synchronize()
        e.set_stream(None)
        assert torch.allclose(Ccm.t().double(), ref, rtol=0, atol=2e-6 * k)


# ---- QR ---------------------------------------------------------------------------------------
def test_qr_kats(L):  # src/qr.rs:257-276
    a = np.array([[3.2, 1.3], [4.4, 5.2], [1.3, 6.7]])
    q, r = L.qr(a).into_decomp()
    np.testing.assert_allclose(q, [[0.5720674, -0.4115578], [0.7865927, 0.0301901], [0.2324024, 0.9108835]], atol=1e-5)
    np.testing.assert_allclose(r, [[5.594, 6.391], [0.0, 5.725]], atol=1e-3)
    q, r = L.qr(np.zeros((2, 2))).into_decomp()
    np.testing.assert_array_equal(q, np.eye(2))
    np.testing.assert_array_equal(r, np.zeros((2, 2)))
    q, r = L.qr_into(np.zeros((0, 0))).into_decomp()  # qr.rs:383-388
    assert q.size == 0 and r.size == 0
    with pytest.raises(L.NotThin):
        L.qr_into(np.zeros((2, 3)))
    with pytest.raises(L.NotSquare):
        L.qr_into(np.zeros((3, 2))).inverse()


SHAPES = [(1, 1), (2, 2), (3, 2), (10, 10), (10, 3), (33, 17), (64, 64), (100, 37), (129, 128), (200, 200), (300, 131),
          (513, 260), (640, 512)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", SHAPES)
def test_qr_parity(L, shape, dt):
    a0 = rnd(shape, dt, seed=shape[0] * 1000 + shape[1])
    ref = a0.copy()
    dref = O.qr(ref)
    for name, a in layouts(a0):
        dec = L.qr_into(a)
        t = tol(a0, c=16)
        assert np.max(np.abs(a - ref)) <= t, (name, np.max(np.abs(a - ref)), t)
        assert np.max(np.abs(dec.diag - dref)) <= t, name
    dec = L.qr(a0)
    q, r = dec.into_decomp()
    n = shape[1]
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    e = EPS[dt]
    assert np.linalg.norm(q.T.astype(np.float64) @ q - np.eye(n)) <= 16 * max(shape) * e
    assert np.linalg.norm(q.astype(np.float64) @ r - a0) <= 16 * max(shape) * e * np.linalg.norm(a0)
    qo = O.generate_q(ref, dref)
    assert np.max(np.abs(q - qo)) <= 64 * max(shape) * e


def test_qr_edge_cases(L):
    # zero column in the middle -> None reflection, diag 0
    a0 = rnd((40, 20), seed=5)
    a0[:, 7] = 0.0
    a0[:, 0] = 0.0
    ref = a0.copy(); dref = O.qr(ref)
    a = a0.copy(); dec = L.qr_into(a)
    assert dec.diag[0] == 0.0 and dref[0] == 0.0
    assert np.max(np.abs(a - ref)) <= tol(a0, 16) and np.max(np.abs(dec.diag - dref)) <= tol(a0, 16)
    assert not dec.is_invertible()
    # a None column stores +0.0 (householder.rs:50 `rn.unwrap_or(A::zero())`): the consumers take signum(diag[i])
    # (householder.rs:89, qr.rs:116) and Rust's signum(-0.0) = -1 would flip every later column of Q
    assert np.array_equal(np.signbit(dec.diag), np.signbit(dref))
    for seed in range(4):
        z0 = rnd((60, 12), seed=100 + seed)
        z0[:, 3 + seed] = 0.0
        q, r = L.qr(z0).into_decomp()
        assert np.linalg.norm(q @ r - z0) <= tol(z0, 16)
        assert np.linalg.norm(q.T @ q - np.eye(12)) <= 64 * 60 * EPS[np.float64]
        qo = O.generate_q(*(lambda w: (w, O.qr(w)))(z0.copy()))
        assert np.max(np.abs(q - qo)) <= 64 * 60 * EPS[np.float64]
    # negative zero / negative pivots and a square matrix (length-1 last reflector)
    a0 = -np.abs(rnd((50, 50), seed=6))
    ref = a0.copy(); dref = O.qr(ref)
    a = a0.copy(); dec = L.qr_into(a)
    assert np.max(np.abs(a - ref)) <= tol(a0, 16)
    assert np.array_equal(np.sign(dec.diag), np.sign(dref))
    # tiny entries (tests/qr.rs:55-75)
    inv = L.qr_into(np.eye(5) * 1e-20).inverse()
    np.testing.assert_allclose(inv, np.eye(5) * 1e20, atol=1e-3)


@pytest.mark.parametrize("shape,k", [((2, 2), 3), ((3, 2), 3), ((60, 60), 7), ((200, 150), 33), ((300, 300), 300)])
def test_qr_solve_qtmul_inverse(L, shape, k):
    a0 = rnd(shape, seed=11)
    x = rnd((shape[1], k), seed=12)
    b = a0 @ x
    dec = L.qr(a0)
    bb = b.copy(); dec.qt_mul(bb)
    ref = a0.copy(); dref = O.qr(ref); bo = b.copy(); O.qt_mul(ref, dref, bo)
    assert np.max(np.abs(bb - bo)) <= tol(b, 32)
    sol = dec.solve(b)
    assert np.max(np.abs(sol - x)) <= 1e-7 * np.max(np.abs(x)) * max(shape)
    if shape[0] == shape[1]:
        inv = dec.inverse()
        assert np.max(np.abs(a0 @ inv - np.eye(shape[0]))) <= 1e-7
    # wide least squares (qr.rs:220-227)
    aw = rnd((shape[1], shape[0]), seed=13)
    if aw.shape[0] < aw.shape[1]:
        xw = rnd((aw.shape[1], 2), seed=14)
        bw = aw @ xw
        sw = L.least_squares(aw.copy(), bw)
        assert np.max(np.abs(aw @ sw - bw)) <= 1e-7 * np.max(np.abs(bw)) * max(shape)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape,k", [((5, 3), 2), ((90, 40), 5), ((40, 90), 3), ((257, 257), 4)])
def test_least_squares_fused_vs_oracle_composition(L, shape, k, dt):
    """lfb_least_squares (one call) against the reference's own composition restated with the oracle:
    qr.rs:207-229 = qr_into + solve_into (qt_mul, then the upper solve with |diag|) / solve_tr_into for wide input."""
    a0 = rnd(shape, dt, seed=shape[0] + 7 * shape[1], lo=-1, hi=1)
    b0 = rnd((shape[0], k), dt, seed=k + shape[0], lo=-1, hi=1)
    m, n = shape
    if m >= n:
        ref = a0.copy(); dref = O.qr(ref); bo = b0.copy(); O.qt_mul(ref, dref, bo)
        xo = bo[:n, :].copy(); O.solve_triangular(ref[:n, :n], xo, O.UPPER, ext_diag=np.abs(dref))
    else:
        xo = None      # wide: checked through the minimum-norm property below (x = Q R^-T b solves A x = b exactly)
    for name, a in layouts(a0, square_t=False)[:3]:
        x = L.least_squares_into(a, b0.copy())
        assert x.shape == (n, k)
        if xo is not None:
            assert np.max(np.abs(x - xo)) <= 64 * max(shape) * EPS[dt] * max(1.0, np.max(np.abs(xo))) * np.linalg.cond(a0.astype(np.float64)), name
        a64, x64, b64 = a0.astype(np.float64), x.astype(np.float64), b0.astype(np.float64)
        # normal equations hold for the least-squares / minimum-norm solution
        lhs = a64.T @ (a64 @ x64 - b64) if m >= n else a64 @ x64 - b64
        assert np.linalg.norm(lhs) <= 256 * max(shape) * EPS[dt] * np.linalg.norm(a64) * max(np.linalg.norm(b64), np.linalg.norm(a64) * np.linalg.norm(x64)), name


def test_qr_errors(L):  # src/qr.rs:338-361
    with pytest.raises(L.NonInvertible):
        L.qr(np.zeros((2, 2))).inverse()
    with pytest.raises(L.NonInvertible):
        L.least_squares_into(np.zeros((2, 2)), np.zeros((2, 2)))
    with pytest.raises(L.NonInvertible):
        L.least_squares_into(np.zeros((2, 3)), np.zeros((2, 2)))
    with pytest.raises(L.WrongRows):
        L.qr(np.eye(3)).solve(np.zeros((2, 2)))


# ---- Cholesky -----------------------------------------------------------------------------------
def spd(n, dt=np.float64, seed=0):  # tests/cholesky.rs:9-19
    g = rnd((n, n), np.float64, seed)
    return (g.T @ g + np.eye(n)).astype(dt)


def test_cholesky_kats(L):  # src/cholesky.rs:209-245, tests/cholesky.rs:87-95
    for dt in (np.float64, np.float32):
        arr = np.array([[25.0, 15, -5], [15, 18, 0], [-5, 0, 11]], dtype=dt)
        ch = L.cholesky(arr)
        np.testing.assert_allclose(ch, [[5.0, 0, 0], [3, 3, 0], [-1, 1, 3]], atol=1e-6)
    with pytest.raises(L.NotSquare):
        L.cholesky(np.array([[1.0, 2, 3], [3, 4, 5]]))
    with pytest.raises(L.NotPositiveDefinite):
        L.cholesky(np.array([[1.0, 2], [2, 1]]))
    with pytest.raises(L.NotPositiveDefinite):
        L.cholesky(np.zeros((2, 2)))
    assert L.cholesky(np.zeros((0, 0))).shape == (0, 0)
    assert L.cholesky(np.ones((1, 1)))[0, 0] == 1.0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 10, 63, 64, 65, 100, 128, 129, 200, 300, 515, 1030])
def test_cholesky_parity(L, n, dt):
    a0 = spd(n, dt, seed=n) if dt == np.float64 else (spd(n, np.float64, seed=n) / 1e4 + np.eye(n)).astype(dt)
    ref = a0.copy(); st, _ = O.cholesky(ref, clean=True)
    assert st == 0
    t = tol(a0, 8) if dt == np.float64 else tol(a0, 64)
    for name, a in layouts(a0):
        got = L.cholesky_inplace(a)
        assert np.max(np.abs(got - ref)) <= t, (name, np.max(np.abs(got - ref)), t)
        assert np.all(np.triu(np.asarray(got), 1) == 0)
    d = a0.copy(); L.cholesky_inplace_dirty(d)
    np.testing.assert_array_equal(np.triu(d, 1), np.triu(a0, 1))   # upper untouched (cholesky.rs:17-19)
    assert np.max(np.abs(np.tril(d) - ref)) <= t
    l64 = np.tril(d).astype(np.float64)
    assert np.linalg.norm(l64 @ l64.T - a0) <= 8 * n * EPS[dt] * np.linalg.norm(a0) * (1 if dt == np.float64 else 8)


@pytest.mark.parametrize("n,bad", [(5, 3), (70, 0), (130, 64), (200, 199), (300, 150)])
def test_cholesky_not_positive_definite(L, n, bad):
    a = spd(n, seed=n + 1)
    a[bad, bad] = -1.0
    ref = a.copy(); st, fi = O.cholesky(ref)
    assert st == 1
    with pytest.raises(L.NotPositiveDefinite) as ei:
        L.cholesky_inplace(a.copy())
    assert ei.value.index == fi == bad


@pytest.mark.parametrize("n,k", [(3, 4), (64, 5), (150, 150), (260, 31)])
def test_solvec_invc(L, n, k):
    a = spd(n, seed=3 * n)
    x = rnd((n, k), seed=n + 5)
    b = a @ x
    out = L.solvec(a.copy(), b)
    assert np.max(np.abs(out - x)) <= 1e-5 * max(1.0, np.max(np.abs(x)))      # tests/cholesky.rs:56-62
    ao = a.copy(); bo = b.copy()
    O.cholesky(ao, clean=False); O.solve_triangular(ao, bo, O.LOWER); O.solve_triangular(ao.T, bo, O.UPPER)
    assert np.max(np.abs(out - bo)) <= 1e-6 * max(1.0, np.max(np.abs(x)))
    if n <= 150:
        a2 = (a / np.linalg.norm(a)) + np.eye(n)
        inv = L.invc(a2)
        assert np.max(np.abs(a2 @ inv - np.eye(n))) <= 1e-7                    # tests/cholesky.rs:64-67


# ---- triangular ----------------------------------------------------------------------------------
@pytest.mark.parametrize("uplo", [0, 1])
@pytest.mark.parametrize("n,k", [(1, 1), (2, 3), (9, 10), (64, 64), (65, 7), (130, 200), (300, 129), (520, 3)])
def test_solve_triangular_parity(L, uplo, n, k):
    a0 = rnd((n, n), seed=n * 31 + uplo)
    d = np.diag(a0).copy(); d[np.abs(d) < 1.0] = 1.0                           # tests/triangular.rs:9-21
    # keep the system well conditioned at n >> 10 (random triangular systems are exponentially
    # ill-conditioned): diagonally dominant rows
    a0[np.arange(n), np.arange(n)] = np.sign(d) * (np.abs(d) + 60.0 * n)
    tri = np.triu(a0) if uplo == L.UPPER else np.tril(a0)
    x = rnd((n, k), seed=n + k)
    b0 = tri @ x
    bo = b0.copy(); O.solve_triangular(a0, bo, uplo)                            # reads only the triangle
    scale = np.max(np.abs(bo)) + 1.0
    for (na, a) in layouts(a0):
        for (nb, b) in layouts(b0, square_t=False)[:3]:
            got = L.solve_triangular_inplace(a, b, uplo)
            assert np.max(np.abs(got - bo)) <= 1e-9 * scale * n, (na, nb)
    ext = np.abs(d) + 3.0
    bo = b0.copy(); O.solve_triangular(a0, bo, uplo, ext_diag=ext)
    import linfa_linalg_b200 as LL
    got = LL._solve_tri(L.engine(), a0, b0.copy(), uplo, ext)
    assert np.max(np.abs(got - bo)) <= 1e-9 * (np.max(np.abs(bo)) + 1) * n


def test_triangular_kats(L):  # src/triangular.rs:233-299, tests/triangular.rs:49-180
    from golden_vectors import TRI_KNOWN_A, TRI_KNOWN_X
    a, x = np.array(TRI_KNOWN_A), np.array(TRI_KNOWN_X)
    out = L.solve_triangular(a, a @ x, L.UPPER)
    np.testing.assert_allclose(out, x, atol=1e-4)
    with np.errstate(all="ignore"):
        L.solve_triangular(np.array([[0.0, 3], [2, 0]]), np.zeros((2, 2)), L.LOWER)   # zero diagonal: no crash
    assert L.solve_triangular(np.zeros((0, 0)), np.zeros((0, 0)), L.UPPER).shape == (0, 0)
    with pytest.raises(L.NotSquare):
        L.solve_triangular(np.array([[1.2, 3.3]]), np.array([[1.2, 3.3]]), L.LOWER)
    with pytest.raises(L.WrongRows):
        L.solve_triangular(np.array([[1.1, 2.2], [3.3, 2.1]]), np.array([[2.2, 3.3]]), L.UPPER)
    sq = np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]])
    np.testing.assert_array_equal(L.into_triangular(sq.copy(), L.UPPER), [[1, 2, 3], [0, 5, 6], [0, 0, 9]])
    np.testing.assert_array_equal(L.into_triangular(sq.copy(), L.LOWER), [[1, 0, 0], [4, 5, 0], [7, 8, 9]])


# ---- tridiagonal ---------------------------------------------------------------------------------
def test_tridiagonal_kat(L):  # src/tridiagonal.rs:124-152
    arr = np.array([[4.0, 1, -2, 2], [1, 2, 0, 1], [-2, 0, 3, -2], [2, 1, -2, -1]])
    dec = L.sym_tridiagonal(arr.copy())
    diag, off = dec.into_diagonals()
    np.testing.assert_allclose(diag, [4, 10 / 3, -33 / 25, 149 / 75], atol=1e-5)
    np.testing.assert_allclose(off, [3, 5 / 3, 68 / 75], atol=1e-5)
    dec = L.sym_tridiagonal(arr.copy())
    q = dec.generate_q(); tri = dec.into_tridiag_matrix()
    np.testing.assert_allclose(q @ tri @ q.T, arr, atol=1e-9)
    np.testing.assert_allclose(q @ q.T, np.eye(4), atol=1e-9)
    one = L.sym_tridiagonal(np.array([[1.1]]))
    d, o = one.into_diagonals()
    assert d[0] == 1.1 and o.size == 0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [2, 3, 10, 33, 100, 257])
def test_tridiagonal_parity(L, n, dt):
    g = rnd((n, n), np.float64, seed=n, lo=-1, hi=1)
    a0 = ((g + g.T) / 2).astype(dt)
    ref = a0.copy(); offr = O.sym_tridiagonal(ref)
    c = 64
    for name, a in layouts(a0)[:4]:
        dec = L.sym_tridiagonal(a)
        assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= tol(a0, c), name   # diag + reflectors (lower part)
        assert np.max(np.abs(dec.off_diagonal - offr)) <= tol(a0, c), name
    dec = L.sym_tridiagonal(a0.copy())
    q = dec.generate_q().astype(np.float64); t = dec.into_tridiag_matrix().astype(np.float64)
    assert np.all(np.abs(np.triu(t, 2)) == 0) and np.all(np.abs(np.tril(t, -2)) == 0)
    assert np.linalg.norm(q @ t @ q.T - a0) <= c * n * EPS[dt] * np.linalg.norm(a0)
    assert np.linalg.norm(q @ q.T - np.eye(n)) <= c * n * EPS[dt]
    L.sym_tridiagonal(rnd((n, n), dt, seed=1)).generate_q()                    # non-symmetric must not crash


# ---- eigh (tridiagonalisation + Givens phase, src/eigh.rs) -----------------------------------------
def test_eigh_kats(L):  # src/eigh.rs:357-409, tests/eigh.rs:59-65
    vals = L.eigvalsh(np.array([[6.0, 2], [2, 6]]))
    np.testing.assert_allclose(vals, [8, 4], atol=1e-12)
    vals = L.eigvalsh(np.array([[1.0, -5, 7], [-5, 2, -9], [7, -9, 3]]))
    np.testing.assert_allclose(L.sort_eig_asc(vals), [-6.86819, -3.41558, 16.28378], atol=1e-5)
    for a, exp in (([[3.0, 1, 1], [1, 3, 1], [1, 1, 3]], [5, 2, 2]), ([[6.0, 2], [2, 6]], [8, 4]),
                   ([[1.0, -5, 7], [-5, 2, -9], [7, -9, 3]], [16.28378, -3.41558, -6.86819])):
        a = np.array(a)
        vals, vecs = L.sort_eig_desc(L.eigh(a))
        np.testing.assert_allclose(vals, exp, atol=1e-5)
        np.testing.assert_allclose(vecs.T @ vecs, np.eye(len(exp)), atol=1e-5)
        np.testing.assert_allclose(a @ vecs, vecs * vals[None, :], atol=1e-5)
    v32 = L.eigvalsh(np.array([[1, -5, 7], [-5, 2, -9], [7, -9, 3]], dtype=np.float32))
    np.testing.assert_allclose(v32, [16.28378, -3.41558, -6.86819], atol=1e-5)      # unsorted, as the reference's test
    vals, vecs = L.eigh(np.array([[2.5]]))
    assert vals[0] == 2.5 and vecs[0, 0] == 1.0
    vals, vecs = L.eigh(np.zeros((3, 3)))                                             # amax == 0 (eigh.rs:32)
    assert np.all(vals == 0) and np.allclose(np.abs(vecs), np.eye(3))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [2, 3, 10, 33, 100, 257, 600])
def test_eigh_parity(L, n, dt):
    """Against the oracle's symmetric_eig (same order of eigenvalues: both follow eigh.rs:51-128; a rounding-level
    difference of the tridiagonal may only reorder deflation, so the elementwise check is on sorted values) and
    the reference's own properties (tests/eigh.rs:10-27): A v = lambda v, orthogonality, eigvalsh == eigh values."""
    g = rnd((n, n), np.float64, seed=n, lo=-100, hi=100)
    a0 = ((g + g.T) / 2).astype(dt)
    vr, qr_ = O.symmetric_eig(a0.copy(), vectors=True)
    scale = np.linalg.norm(a0.astype(np.float64), 2)
    c = 64
    for name, a in layouts(a0)[:5]:
        vals, vecs = L.eigh_into(a)
        assert np.max(np.abs(np.sort(vals) - np.sort(vr))) <= c * n * EPS[dt] * scale, name
        v64, q64 = vals.astype(np.float64), vecs.astype(np.float64)
        assert np.linalg.norm(a0.astype(np.float64) @ q64 - q64 * v64[None, :]) <= c * n * EPS[dt] * scale, name
        assert np.linalg.norm(q64.T @ q64 - np.eye(n)) <= c * n * EPS[dt], name
    vals_only = L.eigvalsh(a0)
    assert np.max(np.abs(np.sort(vals_only) - np.sort(vr))) <= c * n * EPS[dt] * scale
    if n <= 33:   # same deflation order in practice: unsorted elementwise agreement
        vals, _ = L.eigh(a0)
        assert np.max(np.abs(vals - vr)) <= c * n * EPS[dt] * scale
    L.eigh(rnd((n, n), dt, seed=2))                                                  # non-symmetric must not crash (tests/eigh.rs:52-56)


# ---- svd (bidiagonalisation + Givens phase, src/svd.rs) --------------------------------------------
def test_svd_kats(L):  # src/svd.rs:537-614, tests/svd.rs:66-72
    u, s, vt = L.sort_svd_desc(L.svd(np.array([[3.0, 0], [0, -2]]), True, True))
    np.testing.assert_allclose(s, [3, 2], atol=1e-7)
    np.testing.assert_allclose(u, [[1, 0], [0, -1]], atol=1e-7)
    np.testing.assert_allclose(vt, [[1, 0], [0, 1]], atol=1e-7)
    u, s, vt = L.sort_svd_asc(L.svd(np.array([[3.0, 0], [0, -2]]), True, True))
    np.testing.assert_allclose(s, [2, 3], atol=1e-7)
    np.testing.assert_allclose(u, [[0, 1], [-1, 0]], atol=1e-7)
    np.testing.assert_allclose(vt, [[0, 1], [1, 0]], atol=1e-7)
    u, s, vt = L.svd(np.array([[1.0, 0, -1], [-2, 1, 4]]), False, False)
    np.testing.assert_allclose(s, [0.51371, 4.76824], atol=1e-5)
    assert u is None and vt is None
    big = np.array([[10.74785316637712, -5.994983325167452, -6.064492921857296],
                    [-4.149751381521569, 20.654504205822462, -4.470436210703133],
                    [-22.772715014220207, -1.4554372570788008, 18.108113992170573]]).T
    for a, exp in ((np.array([[-2.0, 1, 4]]), [np.sqrt(21)]), (np.array([[1.0, 1], [1, 1]]), [2, 0]),
                   (np.array([[-3.0, 4], [4.3, 2.1], [6.6, 8.7]]), [11.80876, 5.2633658]),
                   (big, [3.16188022e+01, 2.23811978e+01, 0])):
        u, s, vt = L.sort_svd_desc(L.svd(a, True, True))
        np.testing.assert_allclose(s, exp, atol=1e-5)
        assert not np.any(np.signbit(s))
        np.testing.assert_allclose(u @ np.diag(s) @ vt, a, atol=1e-5)
        for cu, cv in ((False, True), (True, False), (False, False)):
            u2, s2, vt2 = L.sort_svd_desc(L.svd(a, cu, cv))
            assert (u2 is None) == (not cu) and (vt2 is None) == (not cv)
            np.testing.assert_allclose(s2, s, atol=1e-9)
            if cu:
                np.testing.assert_allclose(u2, u, atol=1e-9)
            if cv:
                np.testing.assert_allclose(vt2, vt, atol=1e-9)
    u, s, vt = L.svd(np.array([[0.0]]), True, True)
    assert s[0] == 0 and u[0, 0] == 1 and vt[0, 0] == 1
    u, s, vt = L.svd(np.array([[3, 0], [0, -2]], dtype=np.float32), True, True)
    np.testing.assert_allclose(s, [3, 2], atol=1e-7)
    np.testing.assert_allclose(u, [[1, 0], [0, -1]], atol=1e-7)
    np.testing.assert_allclose(vt, [[1, 0], [0, 1]], atol=1e-7)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (3, 4), (4, 3), (10, 10), (1, 7), (7, 1), (40, 23), (23, 40), (130, 70), (64, 200), (300, 300)])
def test_svd_parity(L, shape, dt):
    """Singular values against the oracle's svd (svd.rs:17-221, same algorithm and order up to rounding-level
    reordering of deflations -> compared sorted) and the reference's own properties (tests/svd.rs:10-43)."""
    a0 = rnd(shape, dt, seed=shape[0] * 1000 + shape[1])
    _, sr, _ = O.svd(a0.copy(), False, False)
    md = min(shape)
    scale = np.linalg.norm(a0.astype(np.float64), 2)
    c = 64
    t = c * max(shape) * EPS[dt] * scale
    for name, a in layouts(a0)[:5]:
        u, s, vt = L.svd_into(a, True, True)
        assert u.shape == (shape[0], md) and vt.shape == (md, shape[1]) and s.shape == (md,)
        assert not np.any(np.signbit(s)), name
        assert np.max(np.abs(np.sort(s) - np.sort(sr))) <= t, name
        u64, s64, vt64 = u.astype(np.float64), s.astype(np.float64), vt.astype(np.float64)
        e = c * max(shape) * EPS[dt]
        assert np.linalg.norm((u64.T @ u64 if shape[0] >= shape[1] else u64 @ u64.T) - np.eye(md)) <= e, name
        assert np.linalg.norm(vt64 @ vt64.T - np.eye(md)) <= e, name
        assert np.linalg.norm(u64 @ np.diag(s64) @ vt64 - a0.astype(np.float64)) <= e * np.linalg.norm(a0.astype(np.float64)), name
    for cu, cv in ((False, True), (True, False), (False, False)):
        u2, s2, vt2 = L.svd(a0, cu, cv)
        assert (u2 is None) == (not cu) and (vt2 is None) == (not cv)
        assert np.max(np.abs(np.sort(s2) - np.sort(sr))) <= t


def test_svd_rank_deficient(L):
    """Vanishing diagonal entries take the cancel_horizontal / cancel_vertical paths (svd.rs:293-370).  On numerically
    rank-deficient input the reference's algorithm itself is only as accurate as its absolute `<= eps` tests allow: the
    oracle (verbatim restatement) deviates from LAPACK by up to ~5e-3 when such an input is perturbed by 1e-15.  The
    bar for the engine is therefore the reference's own accuracy envelope on the same matrix, measured here."""
    rng = np.random.default_rng(5)
    for shape, rank in (((12, 8), 3), ((8, 12), 3), ((20, 20), 7), ((6, 6), 0)):
        a = rng.uniform(-1, 1, (shape[0], rank)) @ rng.uniform(-1, 1, (rank, shape[1]))
        envelope = 1e-10
        for _ in range(16):
            p = a * (1 + 1e-15 * rng.standard_normal(shape))
            _, sp, _ = O.svd(p.copy(), False, False)
            envelope = max(envelope, 10 * np.max(np.abs(np.sort(sp)[::-1] - np.linalg.svd(p, compute_uv=False))))
        u, s, vt = L.svd(a, True, True)
        assert u.shape == (shape[0], min(shape)) and vt.shape == (min(shape), shape[1])
        assert not np.any(np.signbit(s))
        assert np.max(np.abs(np.sort(s)[::-1] - np.linalg.svd(a, compute_uv=False))) <= envelope
        assert np.linalg.norm(u @ np.diag(s) @ vt - a) <= 10 * envelope


# ---- bidiagonal ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (3, 4), (4, 3), (10, 10), (1, 7), (7, 1), (40, 23), (23, 40), (130, 70), (64, 200)])
def test_bidiagonal_parity(L, shape, dt):
    a0 = rnd(shape, dt, seed=shape[0] * 100 + shape[1], lo=-1, hi=1)
    ref = a0.copy(); dr, er = O.bidiagonal(ref)
    c = 64
    for name, a in layouts(a0)[:4]:
        dec = L.bidiagonal(a)
        assert np.max(np.abs(a - ref)) <= tol(a0, c), name
        assert np.max(np.abs(dec.diagonal - dr)) <= tol(a0, c), name
        if er.size:
            assert np.max(np.abs(dec.off_diagonal - er)) <= tol(a0, c), name
    dec = L.bidiagonal(a0.copy())
    u = dec.generate_u().astype(np.float64); vt = dec.generate_vt().astype(np.float64); b = dec.into_b().astype(np.float64)
    md = min(shape)
    assert u.shape == (shape[0], md) and vt.shape == (md, shape[1]) and b.shape == (md, md)
    e = EPS[dt] * c * max(shape)
    assert np.linalg.norm((u.T @ u if shape[0] >= shape[1] else u @ u.T) - np.eye(md)) <= e
    assert np.linalg.norm(vt @ vt.T - np.eye(md)) <= e
    assert np.linalg.norm(u @ b @ vt - a0) <= e * np.linalg.norm(a0)
    assert np.all(b >= 0)


# ---- batched -------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("batch,m,n", [(1, 32, 32), (1000, 32, 32), (77, 16, 5), (300, 8, 8), (5, 32, 1), (64, 20, 20)])
def test_qr_batched_parity(L, batch, m, n, dt):
    a0 = rnd((batch, m, n), dt, seed=batch + m + n, lo=-1, hi=1)
    if batch > 3:
        a0[1, :, 0] = 0          # a None column
        a0[2] = 0                # an all-zero matrix
        if n > 2:
            a0[3:, :, 2] = 0     # None columns AFTER live ones: the stored pivot is +0.0 whatever the running sign is
    ref = a0.copy(); dref = O.qr_batched(ref)
    a = a0.copy(); d = L.qr_batched(a)
    t = 16 * m * EPS[dt] * np.sqrt(m)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(d - dref)) <= t
    assert np.array_equal(d == 0, dref == 0)
    assert np.array_equal(np.signbit(d), np.signbit(dref))       # householder.rs:50: None -> +0.0 (signum(-0.0) = -1 in Rust)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("batch,n", [(1, 32), (1000, 32), (77, 5), (300, 8), (5, 1), (64, 20)])
def test_cholesky_batched_parity(L, batch, n, dt):
    """Batched small Cholesky against the oracle's loop over cholesky.rs:51-83: L elementwise, the strict upper triangle
    untouched (dirty) or zero (clean), and the first failing matrix / row reported like a sequential caller would see it."""
    g = rnd((batch, n, n), np.float64, seed=batch * 3 + n, lo=-1, hi=1)
    a0 = (g @ g.transpose(0, 2, 1) + n * np.eye(n)[None]).astype(dt)
    a0[:, np.triu_indices(n, 1)[0], np.triu_indices(n, 1)[1]] = 7.5          # garbage in the (never read) upper triangles
    for clean in (False, True):
        ref = a0.copy(); fm, fi = O.cholesky_batched(ref, clean)
        assert fm == -1
        a = a0.copy(); L.cholesky_batched(a, clean)
        t = 16 * n * EPS[dt] * np.sqrt(n) * np.max(np.abs(a0))
        assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= t
        iu = np.triu_indices(n, 1)
        assert np.all(a[:, iu[0], iu[1]] == (0 if clean else 7.5))
    if batch > 3 and n > 2:
        bad = a0.copy()
        bad[2, n - 1, n - 1] = -1.0                     # matrix 2 fails at its last row
        bad[1, 1, 1] = -5.0                             # matrix 1 fails first, at row 1
        refb = bad.copy(); fm, fi = O.cholesky_batched(refb, False)
        with pytest.raises(L.NotPositiveDefinite) as ei:
            L.cholesky_batched(bad.copy(), False)
        assert (ei.value.matrix, ei.value.index) == (fm, fi) == (1, 1)


# ---- mid-size properties (no oracle: size-independent checks) -------------------------------------
def test_qr_2048_properties(L):
    m, n = 2304, 2048
    a0 = rnd((m, n), seed=99, lo=-1, hi=1)
    dec = L.qr(a0)
    q, r = dec.into_decomp()
    assert np.all(np.diag(r) >= 0)
    assert np.linalg.norm(q.T @ q - np.eye(n)) <= 8 * m * EPS[np.float64]
    assert np.linalg.norm(q @ r - a0) <= 8 * m * EPS[np.float64] * np.linalg.norm(a0)
    import scipy.linalg as sl
    rl = sl.qr(a0, mode="r")[0][:n]
    rl = rl * np.sign(np.diag(rl))[:, None]      # LAPACK R with the reference's diag >= 0 convention
    assert np.max(np.abs(r - rl)) <= 8 * m * EPS[np.float64] * np.linalg.norm(a0, 2)


def test_cholesky_4096_properties(L):
    n = 4096
    g = rnd((n, n), seed=7, lo=-1, hi=1)
    a0 = (g + g.T) / 2 + n * np.eye(n)
    l = L.cholesky(a0)
    assert np.linalg.norm(l @ l.T - a0) <= 8 * n * EPS[np.float64] * np.linalg.norm(a0)
    assert np.max(np.abs(l - np.linalg.cholesky(a0))) <= 8 * n * EPS[np.float64] * np.linalg.norm(a0, 2)


# ---- TSQR local stage (device API): chunked tree inside the GPU vs the oracle's R -------------------
@pytest.mark.parametrize("rows,cols,chunk", [(5000, 32, 1024), (40000, 64, 4096), (70000, 256, 16384), (3000, 48, 16384)])
def test_tsqr_local_r(L, rows, cols, chunk):
    import ctypes as C
    import torch
    e = L.Engine(0)
    e.set_option("tsqr_chunk", chunk)
    a0 = rnd((rows, cols), seed=rows + cols, lo=-1, hi=1)
    ref = a0.copy(); dref = O.qr(ref); r_ref = O.qr_into_r(ref, dref)
    A = torch.from_numpy(np.ascontiguousarray(a0.T)).cuda()        # row-major (cols, rows) == column-major rows x cols
    R = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    st = e.call("lfb_tsqr_local_r_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(R.data_ptr()), cols)
    e._check(st)
    torch.cuda.synchronize()
    r = R.t().cpu().numpy()                                          # column-major cols x cols
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    assert np.max(np.abs(r - r_ref)) <= 64 * cols * EPS[np.float64] * np.linalg.norm(a0, 2) * 8
    e.close()
