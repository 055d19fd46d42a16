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


@pytest.mark.parametrize("n", [777, 1100, 2500])
def test_tridiagonal_large_vs_oracle_and_first_generation(n):
    """The cluster head kernel + lower-triangle SYMV path (several 512-row cluster blocks, interior and
    diagonal SYMV tiles, odd/even n) against the oracle (tridiagonal.rs:31-66) and the per-launch path."""
    import linfa_linalg_b200 as L
    import oracle as O
    g = np.random.default_rng(n).uniform(-1, 1, (n, n))
    a0 = (g + g.T) / 2
    t = 64 * n * EPS * np.linalg.norm(a0)
    a2 = a0.copy()
    dec = L.sym_tridiagonal(a2)
    e1 = L.Engine(0)
    e1.set_option("trd_fused", 0)
    a1 = a0.copy()
    dec1 = L.sym_tridiagonal(a1, e1)
    e1.close()
    assert np.max(np.abs(np.tril(a2) - np.tril(a1))) <= t
    assert np.max(np.abs(dec.off_diagonal - dec1.off_diagonal)) <= t
    e3 = L.Engine(0)
    e3.set_option("trd_symv_async", 0)           # register-staged SYMV tiles instead of cp.async
    a3 = a0.copy()
    dec3 = L.sym_tridiagonal(a3, e3)
    e3.close()
    assert np.max(np.abs(np.tril(a3) - np.tril(a1))) <= t
    assert np.max(np.abs(dec3.off_diagonal - dec1.off_diagonal)) <= t
    if n <= 1100:
        ref = a0.copy(); offr = O.sym_tridiagonal(ref)
        assert np.max(np.abs(np.tril(a2) - np.tril(ref))) <= t
        assert np.max(np.abs(dec.off_diagonal - offr)) <= t
    dec = L.sym_tridiagonal(a0.copy())
    q = dec.generate_q(); tri = dec.into_tridiag_matrix()
    assert np.linalg.norm(q @ tri @ q.T - a0) <= 64 * n * EPS * np.linalg.norm(a0)
    assert np.linalg.norm(q @ q.T - np.eye(n)) <= 64 * n * EPS


@pytest.mark.parametrize("shape", [(700, 300), (300, 700), (1500, 1101), (2600, 640)])
def test_bidiagonal_blocked_vs_oracle_and_unblocked(shape):
    """Blocked bidiagonalisation (cluster head kernels, cp.async GEMV tiles, deferred rank-1 updates, several
    panels, odd sizes, the transposed wide case) against the oracle (bidiagonal.rs:27-59) and the one-reflector-
    at-a-time path; the running sign of householder.rs:45-48 must come out elementwise."""
    import linfa_linalg_b200 as L
    import oracle as O
    m, n = shape
    a0 = np.random.default_rng(m * 7 + n).uniform(-1, 1, shape)
    t = 64 * max(shape) * EPS * np.linalg.norm(a0)
    a2 = a0.copy()
    dec = L.bidiagonal(a2)
    e1 = L.Engine(0)
    e1.set_option("bd_blocked", 0)
    a1 = a0.copy()
    dec1 = L.bidiagonal(a1, e1)
    e1.close()
    assert np.max(np.abs(a2 - a1)) <= t
    assert np.max(np.abs(dec.diagonal - dec1.diagonal)) <= t
    assert np.max(np.abs(dec.off_diagonal - dec1.off_diagonal)) <= t
    if max(shape) <= 1500:
        ref = a0.copy(); dr, er = O.bidiagonal(ref)
        assert np.max(np.abs(a2 - ref)) <= t
        assert np.max(np.abs(dec.diagonal - dr)) <= t
        assert np.max(np.abs(dec.off_diagonal - er)) <= t
    dec = L.bidiagonal(a0.copy())
    u = dec.generate_u(); vt = dec.generate_vt(); b = dec.into_b()
    md = min(shape)
    eb = 64 * max(shape) * EPS
    assert np.linalg.norm((u.T @ u if m >= n else u @ u.T) - np.eye(md)) <= eb
    assert np.linalg.norm(vt @ vt.T - np.eye(md)) <= eb
    assert np.linalg.norm(u @ b @ vt - a0) <= eb * np.linalg.norm(a0)
    assert np.all(b >= 0)


@pytest.mark.parametrize("use_stream", [False, True])
def test_tsqr_local_graph_replay(use_stream):
    """The chunked TSQR stage is captured into a CUDA graph on the second call with the same buffers and
    replayed afterwards (legacy NULL stream: fenced onto the handle's own stream); every call must give the
    oracle's R for the data that is in the buffer at that time."""
    import ctypes as C
    import torch
    import linfa_linalg_b200 as L
    import oracle as O
    rows, cols, chunk = 36000, 64, 4096
    e = L.Engine(0)
    e.set_option("tsqr_chunk", chunk)
    e.set_option("tsqr_graph", 1)
    A = torch.empty((cols, rows), dtype=torch.float64, device="cuda")       # column-major rows x cols
    R = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream() if use_stream else torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)
    l_prev = None
    with torch.cuda.stream(stream):
        for it in range(4):
            a0 = np.random.default_rng(100 + it).uniform(-1, 1, (rows, cols))
            A.copy_(torch.from_numpy(np.ascontiguousarray(a0.T)), non_blocking=False)
            l0 = e.launch_count
            st = e.call("lfb_tsqr_local_r_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(R.data_ptr()), cols)
            e._check(st)
            stream.synchronize()
            nl = e.launch_count - l0
            assert nl > 0 and (l_prev is None or nl == l_prev)
            l_prev = nl
            r = R.t().cpu().numpy()
            ref = a0.copy(); dref = O.qr(ref); r_ref = O.qr_into_r(ref, dref)
            assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
            assert np.max(np.abs(r - r_ref)) <= 64 * cols * EPS * np.linalg.norm(a0, 2) * 8
    e.close()


def test_eigh_1500_properties():
    """eigh.rs end to end at a size where the rotation wavefront runs many passes: eigenvalues against LAPACK,
    residual and orthogonality (tests/eigh.rs:10-27)."""
    import linfa_linalg_b200 as L
    n = 1500
    g = np.random.default_rng(11).uniform(-1, 1, (n, n))
    a0 = (g + g.T) / 2
    vals, vecs = L.eigh(a0)
    s = np.linalg.norm(a0, 2)
    assert np.max(np.abs(np.sort(vals) - np.linalg.eigvalsh(a0))) <= 64 * n * EPS * s
    assert np.linalg.norm(a0 @ vecs - vecs * vals[None, :]) <= 64 * n * EPS * s
    assert np.linalg.norm(vecs.T @ vecs - np.eye(n)) <= 64 * n * EPS
    assert np.max(np.abs(np.sort(L.eigvalsh(a0)) - np.sort(vals))) <= 64 * n * EPS * s


@pytest.mark.parametrize("order", ["C", "F"])
def test_cholesky_dirty_pinned_overlapped_download(order):
    """Page-locked host matrices take the overlapped path: every finished block column of L is copied back while the
    trailing updates still run.  Same contract as the plain path: strict upper triangle untouched, L L^T = A."""
    import torch
    import linfa_linalg_b200 as L
    n = 3000
    rng = np.random.default_rng(8)
    g = rng.uniform(-1, 1, (n, n))
    a0 = (g + g.T) / 2 + n * np.eye(n)
    a0[np.triu_indices(n, 1)] = rng.uniform(5, 6, n * (n - 1) // 2)
    pinned = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    a = pinned.numpy() if order == "C" else pinned.numpy().T            # row-major / column-major views of pinned memory
    a[...] = a0
    L.cholesky_inplace_dirty(a)
    np.testing.assert_array_equal(np.triu(a, 1), np.triu(a0, 1))
    l = np.tril(a)
    sym = np.tril(a0) + np.tril(a0, -1).T
    assert np.linalg.norm(l @ l.T - sym) <= 8 * n * EPS * np.linalg.norm(sym)
    e2 = L.Engine(0)
    e2.set_option("chol_overlap_d2h", 0)
    b = a0.copy()
    L.cholesky_inplace_dirty(b, e2)
    e2.close()
    assert np.max(np.abs(np.tril(a) - np.tril(b))) <= 8 * n * EPS * np.linalg.norm(sym, 2)
