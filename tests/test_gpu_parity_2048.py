"""Round-2 parity additions ("what's weak" / next-round item 1):

* oracle parity raised to the n = 2048 sizes SURVEY.md 8(d) asks for (QR 2304x2048, Cholesky 2048, tridiagonal 2048,
  bidiagonal 2048x1024) and the exact BASELINE config C1 shape (QR 512x512 f64, entries in [-100, 100]);
* batched QR / Cholesky parity at batch = 65536 (elementwise + sign bits);
* eigenvectors / singular vectors ELEMENTWISE against the oracle (same recurrence, so even the column signs agree)
  with `eigh_stable_2x2` = 0 (eigh.rs:111 verbatim) AND 1 (the cancellation-free form), each held to the reference's
  own sensitivity envelope measured on the same matrix;
* the C++ host mirror (include/linfa_b200.hpp) compiled and run (examples/qr_kat.cpp).

The oracle takes about 3 minutes of one host core for the 2048-sized cases (it is the reference's unblocked loops).
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
EPS = {np.float64: 2.220446049250313e-16, np.float32: 1.1920929e-07}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()
    return L


def rnd(shape, dt=np.float64, seed=0, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(dt)


# ---- C1: the reference's own cargo-test-sized case, through the QR trait mirror ------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_qr_c1_512x512(L, dt):
    """BASELINE configs[0]: QR of a 512 x 512 random matrix via `QR::qr` (qr.rs:48-63), entries as tests/common.rs:9."""
    n = 512
    a0 = rnd((n, n), dt, seed=512, lo=-100, hi=100)
    ref = a0.copy(); dref = O.qr(ref)
    dec = L.qr(a0)
    t = 16 * n * EPS[dt] * np.linalg.norm(a0.astype(np.float64))
    assert np.max(np.abs(dec.qr - ref)) <= t
    assert np.max(np.abs(dec.diag - dref)) <= t
    assert np.array_equal(np.signbit(dec.diag), np.signbit(dref))
    q, r = dec.into_decomp()
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    q64 = q.astype(np.float64)
    assert np.linalg.norm(q64.T @ q64 - np.eye(n)) <= 16 * n * EPS[dt]
    assert np.linalg.norm(q64 @ r - a0) <= 16 * n * EPS[dt] * np.linalg.norm(a0.astype(np.float64))
    assert np.max(np.abs(q - O.generate_q(ref, dref))) <= 64 * n * EPS[dt]


# ---- n = 2048 against the oracle ----------------------------------------------------------------------------------
def test_qr_2304x2048_vs_oracle(L):
    m, n = 2304, 2048
    a0 = rnd((m, n), seed=2304)
    ref = np.asfortranarray(a0); dref = O.qr(ref)          # column-major: the oracle's column sweeps run ~3x faster; same sums
    a = a0.copy()
    dec = L.qr_into(a)
    t = 16 * m * EPS[np.float64] * np.linalg.norm(a0)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(dec.diag - dref)) <= t
    assert np.array_equal(np.signbit(dec.diag), np.signbit(dref))
    # unit-norm reflectors (householder.rs:23) and the oracle's R on the nose
    v = np.tril(a)
    assert np.max(np.abs(np.sqrt((v * v).sum(axis=0)) - 1)) <= 64 * EPS[np.float64]
    assert np.max(np.abs(dec.into_r() - O.qr_into_r(ref, dref))) <= t


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_cholesky_2048_vs_oracle(L, dt):
    n = 2048
    g = rnd((n, n), np.float64, seed=2048)
    a0 = (g.T @ g + np.eye(n)).astype(dt)                   # the reference's recipe, tests/cholesky.rs:9-19
    ref = a0.copy(); st, _ = O.cholesky(ref, clean=True)
    assert st == 0
    l = L.cholesky(a0)
    assert np.all(np.triu(l, 1) == 0)
    # kappa(A^T A + I) ~ 3e3 here: the oracle and the blocked engine are both backward stable, so hold them to each other
    # at c n eps ||L|| sqrt(kappa) (forward-error bound of a Cholesky factor) and to A at c n eps ||A||
    kap = np.linalg.cond(a0.astype(np.float64))
    assert np.max(np.abs(l - ref)) <= 8 * n * EPS[dt] * np.linalg.norm(ref.astype(np.float64), 2) * np.sqrt(kap)
    l64 = l.astype(np.float64)
    assert np.linalg.norm(l64 @ l64.T - a0) <= 8 * n * EPS[dt] * np.linalg.norm(a0.astype(np.float64))
    # and the well-conditioned synthetic of the bench: elementwise at 8 n eps ||A||_2
    s0 = ((g + g.T) / 2 + n * np.eye(n)).astype(dt)
    ref = s0.copy(); O.cholesky(ref, clean=False)
    w = s0.copy(); L.cholesky_inplace_dirty(w)
    assert np.max(np.abs(np.tril(w) - np.tril(ref))) <= 8 * n * EPS[dt] * np.linalg.norm(s0.astype(np.float64), 2)
    assert np.array_equal(np.triu(w, 1), np.triu(s0, 1))


def test_tridiagonal_2048_vs_oracle(L):
    n = 2048
    g = rnd((n, n), seed=20481)
    a0 = (g + g.T) / 2
    ref = a0.copy(); offr = O.sym_tridiagonal(ref)
    a = a0.copy()
    dec = L.sym_tridiagonal(a)
    t = 64 * n * EPS[np.float64] * np.linalg.norm(a0)
    assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= t
    assert np.max(np.abs(dec.off_diagonal - offr)) <= t
    d, off = dec.into_diagonals()
    assert np.all(off >= 0)
    assert np.max(np.abs(d - np.diag(ref))) <= t


def test_bidiagonal_2048x1024_vs_oracle(L):
    shape = (2048, 1024)
    a0 = rnd(shape, seed=20482)
    ref = a0.copy(); dr, er = O.bidiagonal(ref)
    a = a0.copy()
    dec = L.bidiagonal(a)
    t = 64 * shape[0] * EPS[np.float64] * np.linalg.norm(a0)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(dec.diagonal - dr)) <= t
    assert np.max(np.abs(dec.off_diagonal - er)) <= t
    assert np.array_equal(np.signbit(dec.diagonal), np.signbit(dr))


# ---- batched at batch >= 65536 --------------------------------------------------------------------------------------
def test_qr_batched_65536_vs_oracle(L):
    batch = 65536
    a0 = rnd((batch, 32, 32), np.float32, seed=65536)
    a0[1, :, 0] = 0
    a0[2] = 0
    a0[3::4097, :, 2] = 0
    ref = a0.copy(); dref = O.qr_batched(ref)
    a = a0.copy(); d = L.qr_batched(a)
    t = 16 * 32 * EPS[np.float32] * np.sqrt(32)
    assert np.max(np.abs(a - ref)) <= t
    assert np.max(np.abs(d - dref)) <= t
    assert np.array_equal(np.signbit(d), np.signbit(dref))
    assert np.array_equal(d == 0, dref == 0)


def test_cholesky_batched_65536_vs_oracle(L):
    batch, n = 65536, 32
    g = rnd((batch, n, n), np.float32, seed=65537)
    a0 = (g @ g.transpose(0, 2, 1) + n * np.eye(n, dtype=np.float32)[None]).astype(np.float32)
    ref = a0.copy()
    fm, fi = O.cholesky_batched(ref, False)
    assert fm == -1
    a = a0.copy(); L.cholesky_batched(a, False)
    t = 16 * n * EPS[np.float32] * np.sqrt(n) * np.max(np.abs(a0))
    assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= t
    iu = np.triu_indices(n, 1)
    assert np.array_equal(a[:, iu[0], iu[1]], a0[:, iu[0], iu[1]])


# ---- eigenvectors / singular vectors elementwise ------------------------------------------------------------------
def _separated_sym(n, seed, dt):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.linspace(1.0, n, n) * rng.choice([-1.0, 1.0], n)          # gaps >= 1, ||A|| = n
    a = (q * lam) @ q.T
    return ((a + a.T) / 2).astype(dt)


@pytest.mark.parametrize("stable", [0, 1])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [5, 33, 100, 257])
def test_eigh_vectors_elementwise_vs_oracle(L, n, dt, stable):
    """eigh.rs:10-129 end to end on a well-separated spectrum: eigenvalues in the reference's own (unsorted) order and the
    eigenvector matrix elementwise, column signs included (both sides run the same implicit-QR recurrence, eigh.rs:51-128).
    `eigh_stable_2x2` = 0 is the reference's 2x2 basis verbatim (eigh.rs:103-121), 1 the engine's cancellation-free default;
    the tolerance is the larger of c n^2 eps and 10x the oracle's own response to 1e-15-relative input perturbations."""
    a0 = _separated_sym(n, 1000 + n, dt)
    vr, qr_ = O.symmetric_eig(a0.copy(), vectors=True)
    rng = np.random.default_rng(n)
    env_v, env_q = 0.0, 0.0
    for _ in range(6):
        p = a0.astype(np.float64) * (1 + 4 * EPS[dt] * rng.standard_normal((n, n)))
        p = ((p + p.T) / 2).astype(dt)
        v1, q1 = O.symmetric_eig(p, vectors=True)
        env_v = max(env_v, float(np.max(np.abs(v1 - vr))))
        env_q = max(env_q, float(np.max(np.abs(q1 - qr_))))
    e = L.Engine(0)
    e.set_option("eigh_stable_2x2", stable)
    vals, vecs = L.eigh(a0, eng=e)
    e.close()
    tv = max(64 * n * EPS[dt] * n, 10 * env_v)
    tq = max(64 * n * n * EPS[dt], 10 * env_q)
    if np.max(np.abs(vals - vr)) <= tv:
        # the usual case: same deflation order -> same (unsorted) eigenvalue order, same column signs
        assert np.max(np.abs(vecs - qr_)) <= tq, (np.max(np.abs(vecs - qr_)), tq, env_q)
    else:
        # a rounding-level difference of the tridiagonal may reorder deflations (eigh.rs:60-100 tests `<= eps` quantities): the
        # eigenvalue ORDER then differs from the oracle's run although both follow the reference.  Match by value; a column's
        # sign depends on the rotation sequence, so compare up to sign in that case.
        i0, i1 = np.argsort(vr, kind="stable"), np.argsort(vals, kind="stable")
        assert np.max(np.abs(vals[i1] - vr[i0])) <= tv
        a, b = qr_[:, i0], vecs[:, i1]
        d = np.minimum(np.max(np.abs(a - b), axis=0), np.max(np.abs(a + b), axis=0))
        assert np.max(d) <= tq, (np.max(d), tq, env_q)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(7, 7), (40, 23), (23, 40), (130, 70), (200, 200)])
def test_svd_vectors_elementwise_vs_oracle(L, shape, dt):
    """svd.rs:17-221 on well-separated singular values: sigma in the reference's order, U and Vt elementwise (signs
    included: svd.rs:158-192's 2x2 kernel and the final sign fix :213-221 are the same on both sides)."""
    m, n = shape
    md = min(shape)
    rng = np.random.default_rng(m * 100 + n)
    u0, _ = np.linalg.qr(rng.standard_normal((m, md)))
    v0, _ = np.linalg.qr(rng.standard_normal((n, md)))
    a0 = ((u0 * np.linspace(1.0, md, md)) @ v0.T).astype(dt)
    ur, sr, vtr = O.svd(a0.copy(), True, True)
    env_s, env_u = 0.0, 0.0
    for _ in range(6):
        p = (a0.astype(np.float64) * (1 + 4 * EPS[dt] * rng.standard_normal(shape))).astype(dt)
        u1, s1, vt1 = O.svd(p, True, True)
        env_s = max(env_s, float(np.max(np.abs(s1 - sr))))
        env_u = max(env_u, float(np.max(np.abs(u1 - ur))), float(np.max(np.abs(vt1 - vtr))))
    u, s, vt = L.svd(a0, True, True)
    big = max(shape)
    ts = max(64 * big * EPS[dt] * md, 10 * env_s)
    tu = max(64 * big * md * EPS[dt], 10 * env_u)
    if np.max(np.abs(s - sr)) <= ts:
        assert np.max(np.abs(u - ur)) <= tu, (np.max(np.abs(u - ur)), tu)
        assert np.max(np.abs(vt - vtr)) <= tu, (np.max(np.abs(vt - vtr)), tu)
    else:   # deflations reordered by a rounding-level difference (see the eigh test): match by value, joint sign of (u_k, v_k)
        i0, i1 = np.argsort(sr, kind="stable"), np.argsort(s, kind="stable")
        assert np.max(np.abs(s[i1] - sr[i0])) <= ts
        sg = np.sign(np.sum(ur[:, i0] * u[:, i1], axis=0))
        assert np.max(np.abs(ur[:, i0] - u[:, i1] * sg)) <= tu
        assert np.max(np.abs(vtr[i0] - vt[i1] * sg[:, None])) <= tu


# ---- the C++ mirror --------------------------------------------------------------------------------------------------
def test_cpp_mirror_example_compiles_and_runs(tmp_path):
    """include/linfa_b200.hpp + examples/qr_kat.cpp: the reference's QR / Cholesky / eigh / svd KATs through the C++ host
    mirror, linked against the in-tree liblinfa_b200.so."""
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    libdir = os.path.join(ROOT, "linfa_linalg_b200", "lib")
    exe = str(tmp_path / "qr_kat")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "qr_kat.cpp"), "-L" + libdir, "-llinfa_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max deviation from the reference KATs" in r.stdout
