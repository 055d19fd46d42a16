"""The f32 tensor-core path (csrc/gemm_tf32.cu: tcgen05.mma.kind::tf32, TMEM accumulator, TMA operands, 3xTF32 split).

Stated tolerance of 3xTF32: every product carries a relative error <= 2^-21 (an exact f32 FMA: 2^-24), accumulation is f32.
The GEMM is held to |C - C_f64| <= 2e-6 * K on operands in [-0.5, 0.5] (the bound the FFMA kernel is held to in
test_gpu_parity.py::test_gemm_f32), and the f32 factorisations that run on it to the same c n eps_32 bounds against the f32
oracle as before (`sgemm_tc` = 1, the default) -- with the FFMA kernel (`sgemm_tc` = 0) beside it on the same inputs."""
import ctypes as C

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
EPS32 = 1.1920929e-07


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()
    return L


def _gemm_tn(e, A, B, Cm, alpha, beta):
    """column-major C (m x n) = alpha A^T B + beta C with A stored k x m, B stored k x n (torch tensors hold the storage)."""
    import torch
    m, k = A.shape          # tensor A: (m, k) row-major == column-major k x m
    n = B.shape[0]
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    st = e.call("lfb_gemm_dev_f32", 1, 0, m, n, k, alpha, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0),
                beta, C.c_void_p(Cm.data_ptr()), Cm.stride(0))
    e._check(st)
    torch.cuda.synchronize()
    e.set_stream(None)


@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (128, 128, 64), (256, 384, 1000), (200, 300, 100), (129, 130, 36), (1024, 512, 4096),
                                   (128, 4096, 8192)])
def test_sgemm_tn_tensor_core(L, m, n, k, mode):
    import torch
    e = L.Engine(0)
    e.set_option("sgemm_tc", mode)
    g = torch.Generator(device="cuda").manual_seed(m + 3 * n + 7 * k + mode)
    A = torch.rand((m, k), dtype=torch.float32, device="cuda", generator=g) - 0.5        # column-major k x m
    B = torch.rand((n, k), dtype=torch.float32, device="cuda", generator=g) - 0.5        # column-major k x n
    C0 = torch.rand((n, m), dtype=torch.float32, device="cuda", generator=g) - 0.5       # column-major m x n (tensor holds C^T)
    ref = 1.5 * (A.double() @ B.double().t()) + 0.5 * C0.t().double()                    # m x n
    Cm = C0.clone()
    l0 = e.launch_count
    _gemm_tn(e, A, B, Cm, 1.5, 0.5)
    assert e.launch_count > l0
    err = float((Cm.t().double() - ref).abs().max())
    tol = 2e-6 * k if mode != 2 else 3e-3 * k ** 0.5 + 1e-3
    assert err <= tol, (err, tol)
    Cz = torch.full_like(C0, float("nan"))                                               # beta = 0 must not read C
    _gemm_tn(e, A, B, Cz, 1.0, 0.0)
    assert float((Cz.t().double() - A.double() @ B.double().t()).abs().max()) <= tol
    e.close()


def test_3xtf32_is_at_f32_accuracy_where_single_tf32_is_not(L):
    """The split matters: on the same product single-pass TF32 is ~1e-3 off, 3xTF32 is within a few f32 ulps of FFMA."""
    import torch
    m = n = 256
    k = 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((m, k), dtype=torch.float32, device="cuda", generator=g) + 0.5          # all positive: no cancellation
    B = torch.rand((n, k), dtype=torch.float32, device="cuda", generator=g) + 0.5
    ref = A.double() @ B.double().t()
    errs = {}
    for mode in (0, 1, 2):
        e = L.Engine(0)
        e.set_option("sgemm_tc", mode)
        Cm = torch.zeros((n, m), dtype=torch.float32, device="cuda")
        _gemm_tn(e, A, B, Cm, 1.0, 0.0)
        errs[mode] = float(((Cm.t().double() - ref) / ref).abs().max())
        e.close()
    assert errs[1] <= 8 * EPS32 * 4, errs            # 3xTF32: 2^-21 per product + f32 accumulation
    assert errs[0] <= 8 * EPS32 * 4, errs
    assert errs[2] >= 20 * errs[1], errs             # single TF32 is visibly worse (that is why it is never a default)


@pytest.mark.parametrize("mode", [1, 0])
def test_cholesky_f32_2560_on_tensor_cores(L, mode):
    """cholesky.rs:51-83 in f32 at a size whose trailing SYRKs run on the tcgen05 kernel (TN form, lower-only tiles)."""
    n = 2560
    g = np.random.default_rng(11).uniform(-1, 1, (n, n))
    a0 = ((g + g.T) / 2 + n * np.eye(n)).astype(np.float32)
    a0[np.triu_indices(n, 1)] = 7.5                                   # never read, never written (dirty contract)
    ref = a0.copy(); st, _ = O.cholesky(ref, clean=False)
    assert st == 0
    e = L.Engine(0)
    e.set_option("sgemm_tc", mode)
    a = a0.copy()
    L.cholesky_inplace_dirty(a, eng=e)
    e.close()
    assert np.all(a[np.triu_indices(n, 1)] == 7.5)
    sym = np.tril(a0) + np.tril(a0, -1).T
    assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= 8 * n * EPS32 * np.linalg.norm(sym.astype(np.float64), 2)
    l = np.tril(a).astype(np.float64)
    assert np.linalg.norm(l @ l.T - sym) <= 8 * n * EPS32 * np.linalg.norm(sym.astype(np.float64))


@pytest.mark.parametrize("mode", [1, 0])
def test_qr_f32_1536x1280_on_tensor_cores(L, mode):
    """qr.rs:29-45 in f32: W = V^T C and C -= V W (through the transposed V copy) on the tcgen05 kernel."""
    m, n = 1536, 1280
    a0 = np.random.default_rng(12).uniform(-1, 1, (m, n)).astype(np.float32)
    ref = np.asfortranarray(a0); dref = O.qr(ref)
    e = L.Engine(0)
    e.set_option("sgemm_tc", mode)
    a = a0.copy()
    dec = L.qr_into(a, eng=e)
    t = 16 * m * EPS32 * np.linalg.norm(a0.astype(np.float64))
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(dec.diag - dref)) <= t
    assert np.array_equal(np.signbit(dec.diag), np.signbit(dref))
    q, r = dec.into_decomp()
    e.close()
    q64 = q.astype(np.float64)
    assert np.linalg.norm(q64.T @ q64 - np.eye(n)) <= 16 * m * EPS32
    assert np.linalg.norm(q64 @ r - a0) <= 16 * m * EPS32 * np.linalg.norm(a0.astype(np.float64))


@pytest.mark.parametrize("dt,ta,m,n,k", [("f64", 1, 128, 256, 200000), ("f64", 1, 37, 19, 100001), ("f64", 0, 64, 40, 50000),
                                         ("f32", 1, 128, 128, 65536), ("f32", 1, 33, 70, 100001)])
def test_split_k_is_bit_reproducible(L, dt, ta, m, n, k):
    """Skinny outputs with a long K are split over CTAs; the partial tiles are summed by a reduce kernel in a fixed order
    (`gemm_deterministic`, default), so repeated calls give bit-identical results -- like the reference's sequential loops --
    on all four GEMM kernels (f64 TMA, f64 fallback, f32 tcgen05, f32 FFMA fallback)."""
    import torch
    tdt = torch.float64 if dt == "f64" else torch.float32
    e = L.Engine(0)
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    A = torch.rand((m, k) if ta else (k, m), dtype=tdt, device="cuda", generator=g) - 0.5       # column-major storage of op's operand
    B = torch.rand((n, k), dtype=tdt, device="cuda", generator=g) - 0.5                         # column-major k x n
    outs = []
    for _ in range(3):
        Cm = torch.zeros((n, m), dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        st = e.call("lfb_gemm_dev_" + dt, ta, 0, m, n, k, 1.0, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0),
                    0.0, C.c_void_p(Cm.data_ptr()), m)
        e._check(st)
        torch.cuda.synchronize()
        e.set_stream(None)
        outs.append(Cm.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = (A.double() if ta else A.double().t()) @ B.double().t()
    tol = (1e-11 if dt == "f64" else 2e-6) * k
    assert float((outs[0].t().double() - ref).abs().max()) <= tol
    e.close()
