"""GPU parity of the tall-skinny route of qr_into: TSQR + Householder reconstruction (csrc/tsqr_hr.cu, SURVEY.md 8f
rank 2) must deliver the reference's own QRDecomp -- same unit-norm reflectors, same R rows, same signed pivots
(qr.rs:29-45, householder.rs:9-51 through the oracle) -- and the reference's consumers must accept it.

Tolerances: reflector entries are O(1) (unit-norm vectors): |v - v_oracle| <= 64 * cols * eps absolute; R rows and
diag: <= 64 * cols * eps * ||A||_2; properties as tests/qr.rs:9-53 (Q^T Q = I, Q R = A) at 64 * rows * eps.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

EPS = {np.float64: 2.220446049250313e-16, np.float32: 1.1920929e-07}


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()
    return L


def rnd(shape, dt=np.float64, seed=0, lo=-100.0, hi=100.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(dt)


def check_against_oracle(a0, factor, diag, c=64.0):
    dt = a0.dtype.type
    rows, cols = a0.shape
    ref = a0.copy()
    dref = O.qr(ref)
    eps = EPS[dt]
    nrm = np.linalg.norm(a0.astype(np.float64), 2)
    lo, lo_ref = np.tril(factor), np.tril(ref)
    assert np.max(np.abs(lo - lo_ref)) <= c * cols * eps, "reflectors differ from householder.rs:9-28"
    assert np.max(np.abs(np.triu(factor, 1) - np.triu(ref, 1))) <= c * cols * eps * nrm, "R rows differ"
    assert np.max(np.abs(diag - dref)) <= c * cols * eps * nrm, "signed pivots differ"
    assert np.array_equal(np.signbit(diag), np.signbit(dref)), "pivot signs differ (householder.rs:16)"


@pytest.mark.parametrize("rows,cols,chunk", [(5000, 32, 1024), (40000, 64, 4096), (30011, 128, 4096), (9000, 17, 512),
                                             (3000, 48, 16384), (64, 64, 16384), (1, 1, 16384)])
def test_qr_tsqr_matches_reference_factor_f64(L, rows, cols, chunk):
    e = L.Engine(0)
    e.set_option("tsqr_chunk", chunk)
    a0 = rnd((rows, cols), seed=rows + cols)
    a = a0.copy()
    dec = L.qr_tsqr_into(a, eng=e)
    check_against_oracle(a0, a, dec.diag)
    e.close()


@pytest.mark.parametrize("rows,cols,chunk", [(20000, 32, 2048), (6000, 24, 16384)])
def test_qr_tsqr_matches_reference_factor_f32(L, rows, cols, chunk):
    e = L.Engine(0)
    e.set_option("tsqr_chunk", chunk)
    a0 = rnd((rows, cols), dt=np.float32, seed=rows + cols, lo=-1, hi=1)
    a = a0.copy()
    dec = L.qr_tsqr_into(a, eng=e)
    check_against_oracle(a0, a, dec.diag)
    e.close()


def test_qr_tsqr_layouts_and_consumers(L):
    """Column-major and strided host views; generate_q / into_r / qt_mul / solve_into on the result (qr.rs:86-152)."""
    e = L.Engine(0)
    e.set_option("tsqr_chunk", 1024)
    rows, cols = 7000, 40
    a0 = rnd((rows, cols), seed=5)
    for a in (np.array(a0, order="F"), np.array(a0[::-1], order="C")[::-1]):
        dec = L.qr_tsqr_into(a, eng=e)
        check_against_oracle(a0, a, dec.diag)
    q, r = dec.into_decomp()
    eps = EPS[np.float64]
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    assert np.linalg.norm(q.T @ q - np.eye(cols)) <= 64 * rows * eps
    assert np.linalg.norm(q @ r - a0) <= 64 * rows * eps * np.linalg.norm(a0)
    b = rnd((rows, 3), seed=6)
    x = L.qr_tsqr(a0, eng=e).solve_into(b.copy())
    xr = np.linalg.lstsq(a0, b, rcond=None)[0]
    assert np.max(np.abs(x - xr)) <= 1e-9 * np.max(np.abs(xr))
    e.close()


def test_qr_auto_route(L):
    """With qr_tsqr_auto the plain qr_into takes the TSQR route for tall-skinny inputs: same results either way."""
    e = L.Engine(0)
    e.set_option("tsqr_chunk", 2048)
    a0 = rnd((20000, 48), seed=9)
    a1 = a0.copy()
    d1 = L.qr_into(a1, eng=e).diag
    e.set_option("qr_tsqr_auto", 1)
    a2 = a0.copy()
    d2 = L.qr_into(a2, eng=e).diag
    assert not np.array_equal(a1, a2)           # a different algorithm ran (rounding differs) ...
    check_against_oracle(a0, a1, d1)            # ... to the same factor
    check_against_oracle(a0, a2, d2)
    e.close()


@pytest.mark.parametrize("kind", ["zeros", "zero_col", "dup_col", "diag"])
def test_qr_tsqr_degenerate(L, kind):
    """Rank-deficient input: R is not unique, so the gate is the reference's own properties (tests/qr.rs:9-53) and the
    error behaviour (qr.rs:194-197: NonInvertible iff a pivot is exactly zero)."""
    e = L.Engine(0)
    e.set_option("tsqr_chunk", 256)
    rows, cols = 3000, 6
    a0 = rnd((rows, cols), seed=3, lo=-1, hi=1)
    if kind == "zeros":
        a0[:] = 0
    elif kind == "zero_col":
        a0[:, 2] = 0
    elif kind == "dup_col":
        a0[:, 3] = 2 * a0[:, 1]
    else:
        a0[:] = 0
        a0[0, 0], a0[1, 1], a0[2, 2], a0[3, 3], a0[4, 4], a0[5, 5] = 1, -2, 3, -4, 5, -6
    a = a0.copy()
    dec = L.qr_tsqr_into(a, eng=e)
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(dec.diag))
    if kind == "diag":
        check_against_oracle(a0, a, dec.diag)      # before into_decomp: into_r rewrites the top block in place (qr.rs:91-98)
    q, r = dec.into_decomp()
    eps = EPS[np.float64]
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    assert np.linalg.norm(q.T @ q - np.eye(cols)) <= 64 * rows * eps
    assert np.linalg.norm(q @ r - a0) <= 64 * rows * eps * max(np.linalg.norm(a0), 1.0)
    if kind == "zeros":
        assert np.all(r == 0) and np.allclose(q, np.eye(rows, cols), rtol=0, atol=1e-15)      # qr.rs:272-276
        assert not dec.is_invertible()
    e.close()


def test_tsqr_qr_dist_single_rank_gpu_ops(L):
    """dist.tsqr_qr with the CUDA callables (world 1: no collective): the plumbing the multi-GPU bench uses."""
    import torch
    from linfa_linalg_b200 import dist as D
    e = L.Engine(0)
    e.set_option("tsqr_chunk", 2048)
    rows, cols = 25000, 64
    a0 = rnd((rows, cols), seed=21, lo=-1, hi=1)
    blk = torch.from_numpy(np.ascontiguousarray(a0.T)).cuda()
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    diag, r = D.tsqr_qr(blk, D.GpuTsqrOps(e), cols)
    torch.cuda.synchronize()
    e.set_stream(None)
    check_against_oracle(a0, blk.t().cpu().numpy(), diag.cpu().numpy())
    e.close()


def test_tsqr_explicit_q_blocks(L):
    """The building blocks on their own: explicit Q (orthonormal, Q R = A, diag(R) >= 0) and Q <- Q Qs."""
    import torch
    e = L.Engine(0)
    e.set_option("tsqr_chunk", 1024)
    rows, cols = 9000, 32
    a0 = rnd((rows, cols), seed=33, lo=-1, hi=1)
    A = torch.from_numpy(np.ascontiguousarray(a0.T)).cuda()
    R = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e._check(e.call("lfb_tsqr_explicit_q_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(R.data_ptr()), cols))
    torch.cuda.synchronize()
    q = A.t().cpu().numpy()
    r = R.t().cpu().numpy()
    eps = EPS[np.float64]
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    assert np.linalg.norm(q.T @ q - np.eye(cols)) <= 64 * rows * eps
    assert np.linalg.norm(q @ r - a0) <= 64 * rows * eps * np.linalg.norm(a0)
    g = rnd((cols, cols), seed=34, lo=-1, hi=1)
    G = torch.from_numpy(np.ascontiguousarray(g.T)).cuda()
    e._check(e.call("lfb_tsqr_apply_q_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(G.data_ptr()), cols))
    torch.cuda.synchronize()
    e.set_stream(None)
    assert np.max(np.abs(A.t().cpu().numpy() - q @ g)) <= 64 * cols * eps
    e.close()
