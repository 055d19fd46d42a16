"""The blocked QR's 128-column panel as guarded Cholesky-QR + Householder reconstruction (csrc/panel_hr.cu, householder.cu:
panel_cholqr) and the two-CTA-per-SM FP64 GEMM behind its trailing updates (csrc/gemm_tma.cu: dgemm_tma2_kernel).

* parity with the oracle (qr.rs:29-45, householder.rs:9-28: elementwise factor, signed pivots with their sign bits) on inputs
  whose panels the guard ACCEPTS (well-conditioned random), DECLINES (graded / nearly dependent / exactly dependent columns) and
  a mix of both inside one matrix -- whichever route a panel takes, the result is the reference's;
* the three routes (fused single-CTA panel kernels, the same stages as separate launches, cluster Householder panels) agree;
* the 128 x 64-tile GEMM is bitwise equal to the 128 x 128 one (same k order, same epilogue arithmetic), split-K included.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()
    return L


def _qr_with(L, a0, **opts):
    e = L.Engine(0)
    for k, v in opts.items():
        e.set_option(k, v)
    a = a0.copy()
    dec = L.qr_into(a, eng=e)
    return a, dec.diag.copy()


def _cases():
    rng = np.random.default_rng(77)
    m, n = 900, 512
    well = rng.uniform(-1, 1, (m, n))
    graded = well * np.logspace(0, -9, n)[None, :]                       # column norms over 9 decades: cond of a panel ~1e2..1e3
    near = well.copy()
    near[:, 200] = near[:, 199] + 1e-7 * rng.uniform(-1, 1, m)           # one nearly dependent pair inside panel 1
    dep = well.copy()
    dep[:, 300] = dep[:, 290]                                            # exactly dependent: Cholesky fails or cond = inf
    dep[:, 301] = 0.0                                                    # and a `None` pivot (householder.rs:26)
    return {"well": well, "graded": graded, "near": near, "dependent": dep}


@pytest.mark.parametrize("name", ["well", "graded", "near", "dependent"])
def test_qr_panel_routes_match_oracle(L, name):
    a0 = _cases()[name]
    m, n = a0.shape
    ref = np.asfortranarray(a0)
    dref = O.qr(ref)
    scale = np.linalg.norm(a0)
    outs = {}
    for tag, opts in {"fused": {}, "separate": {"cholqr_fused": 0}, "householder": {"qr_panel_cholqr": 0}}.items():
        a, d = _qr_with(L, a0, **opts)
        outs[tag] = (a, d)
        if name in ("well", "graded"):
            # elementwise parity needs a well-posed factorisation; the column-scaled case is held per column
            cs = np.maximum(np.linalg.norm(a0, axis=0), 1e-300)
            t = 64 * m * EPS
            assert np.max(np.abs(np.triu(a, 1) - np.triu(ref, 1)) / cs[None, :]) <= t, tag
            assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= 4096 * m * EPS, tag            # unit-norm reflectors
            assert np.max(np.abs(d - dref) / cs) <= t, tag
            assert np.array_equal(np.signbit(d), np.signbit(dref)), tag
        else:
            # (nearly) rank-deficient: reflectors past the dependency are not unique to rounding; hold the invariants instead
            q = O.generate_q(np.ascontiguousarray(a), d)
            r = O.qr_into_r(np.ascontiguousarray(a), d)
            assert np.linalg.norm(q.T @ q - np.eye(n)) <= 64 * n * EPS, tag
            assert np.linalg.norm(q @ r - a0) <= 64 * m * EPS * scale, tag
            assert np.all(np.diag(r) >= 0), tag
    if name == "well":
        for tag in ("separate", "householder"):
            assert np.max(np.abs(outs[tag][0] - outs["fused"][0])) <= 64 * m * EPS * scale
            assert np.array_equal(np.signbit(outs[tag][1]), np.signbit(outs["fused"][1]))


@pytest.mark.parametrize("shape", [(700, 200), (1000, 328), (300, 130), (257, 128), (255, 128), (4100, 384)])
def test_qr_panel_ragged_widths_match_oracle(L, shape):
    """Last panels of 72 / 72 / 2 columns (the separate-launch Cholesky-QR route for 64 <= nb < 128, cluster panels below that) and
    panels just above / below the rows >= 2 nb gate."""
    m, n = shape
    a0 = np.random.default_rng(m * 7 + n).uniform(-1, 1, (m, n))
    ref = np.asfortranarray(a0)
    dref = O.qr(ref)
    a, d = _qr_with(L, a0)
    t = 64 * m * EPS * np.linalg.norm(a0)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(d - dref)) <= t
    assert np.array_equal(np.signbit(d), np.signbit(dref))


def test_qr_panel_2048_large_entries(L):
    """QR 2176 x 2048 with entries in [-100, 100] (tests/common.rs:9 scale): seventeen 128-column panels on the fused route."""
    m, n = 2176, 2048
    a0 = np.random.default_rng(5).uniform(-100, 100, (m, n))
    ref = np.asfortranarray(a0)
    dref = O.qr(ref)
    a, d = _qr_with(L, a0)
    t = 16 * m * EPS * np.linalg.norm(a0)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(d - dref)) <= t
    assert np.array_equal(np.signbit(d), np.signbit(dref))


@pytest.mark.parametrize("shape", [(0, 0, 2048, 1920, 128), (1, 0, 2048, 2048, 512), (0, 1, 1990, 2110, 64), (1, 1, 1026, 1538, 258),
                                   (1, 0, 128, 4096, 8192), (0, 0, 2050, 1282, 130)])
def test_gemm_two_cta_tiles_bitwise(L, shape):
    import torch
    ta, tb, M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.rand((M, K) if not ta else (K, M), dtype=torch.float64, device="cuda", generator=g) - 0.5
    B = torch.rand((K, N) if not tb else (N, K), dtype=torch.float64, device="cuda", generator=g) - 0.5
    C0 = torch.rand((N, M), dtype=torch.float64, device="cuda", generator=g) - 0.5      # column-major M x N
    Acm, Bcm = A.t().contiguous(), B.t().contiguous()                                   # column-major buffers, ld = A.shape[0]
    outs = []
    for on in (0, 1, 2):
        e = L.Engine(0)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_option("gemm_tma2", on)
        Cm = C0.clone()
        st = e.lib.lfb_gemm_dev_f64(e.h, ta, tb, M, N, K, -1.0, C.c_void_p(Acm.data_ptr()), A.shape[0], C.c_void_p(Bcm.data_ptr()), B.shape[0],
                                    1.0, C.c_void_p(Cm.data_ptr()), M)
        assert st == 0
        torch.cuda.synchronize()
        outs.append(Cm)
    ref = C0.t() - (A.t() if ta else A) @ (B.t() if tb else B)
    assert float((outs[1].t() - ref).abs().max()) <= 64 * K * EPS
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("shape", [(2304, 2048), (4096, 2176)])
def test_qr_host_overlapped_download_is_identical(L, order, shape):
    """Page-locked host memory, cols >= 2048: block columns leave for the host during the factorisation, after a per-panel pass of
    the reference's running sign (householder.rs:45-50).  Negations only: bit-identical to the download-at-the-end path."""
    import torch
    m, n = shape
    a0 = np.random.default_rng(m + n).uniform(-1, 1, (m, n))
    ref = np.array(a0, order=order)
    e0 = L.Engine(0)
    e0.set_option("qr_overlap_d2h", 0)
    d0 = L.qr_into(ref, eng=e0).diag.copy()
    t = torch.from_numpy(a0.copy()).pin_memory()
    a = t.numpy() if order == "C" else None
    if order == "F":
        tf = torch.empty((n, m), dtype=torch.float64).pin_memory()          # column-major m x n view of a pinned buffer
        a = tf.numpy().T
        a[...] = a0
    d1 = L.qr_into(a, eng=L.Engine(0)).diag.copy()
    assert np.array_equal(a, ref)
    assert np.array_equal(d1, d0) and np.array_equal(np.signbit(d1), np.signbit(d0))


def test_qr_f32_two_level_panel_matches_householder_route(L):
    """f32, 2560 x 2304: 256-column panels (K of the tcgen05 update) as two fused 128-column Cholesky-QR panels with the pair's
    compact-WY factor assembled from T_A, T_B and V_A^T V_B -- against the Householder panels (qr_panel_cholqr = 0) and the
    invariants in f64 (R^T R = A^T A with R from into_r, qr.rs:91-98; unit-norm reflectors, householder.rs:23)."""
    m, n = 2560, 2304
    a0 = np.random.default_rng(23).uniform(-1, 1, (m, n)).astype(np.float32)
    a_h, d_h = _qr_with(L, a0, qr_panel_cholqr=0)
    e = L.Engine(0)
    a_c = a0.copy()
    dec = L.qr_into(a_c, eng=e)
    d_c = dec.diag.copy()
    eps = 1.1920929e-07
    scale = float(np.linalg.norm(a0.astype(np.float64)))
    assert np.array_equal(np.signbit(d_c), np.signbit(d_h))
    assert np.max(np.abs(a_c - a_h)) <= 16 * m * eps * scale / np.sqrt(m)
    v = np.tril(a_c.astype(np.float64))
    assert np.max(np.abs(np.sqrt((v * v).sum(axis=0)) - 1)) <= 64 * eps
    r = dec.into_r().astype(np.float64)
    g = a0.astype(np.float64).T @ a0.astype(np.float64)
    assert np.all(np.diag(r) >= 0)
    assert np.linalg.norm(r.T @ r - g) / np.linalg.norm(g) <= 64 * eps


def test_every_documented_option_is_settable(L):
    """Every tunable of lfb::Options (csrc/common.cuh) is registered in lfb_set_option's table and accepts its own default."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "linfa_linalg_b200", "csrc", "common.cuh")).read()
    body = src[src.index("struct Options"):]
    body = body[:body.index("};")]
    opts = re.findall(r"^\s*int64_t\s+([a-z0-9_]+)\s*=\s*(-?\d+)\s*;", body, flags=re.M)
    assert len(opts) >= 40
    e = L.Engine(0)
    missing = [name for name, dflt in opts if e.lib.lfb_set_option(e.h, name.encode(), int(dflt)) != 0]
    assert not missing, missing
