"""CPU: the mathematics of csrc/tsqr_hr.cu stated in NumPy and held to the oracle's qr (qr.rs:29-45) elementwise --
the statement the CUDA code was written from, kept as a test so the conversion formulas stay pinned:

  chunked TSQR with explicit Q  ->  LU without pivoting of Q - [S; 0]  ->  reflectors c_k y_k, diag_k = s_{k-1} s_k R_kk.

Also the variant DESIGN.md section 12 item 6 plans (the right-hand solve folded into the combine GEMM), so that the
next kernel generation starts from a checked formula.
"""
import numpy as np
import pytest

import oracle as O


def local_q_r(a):
    f = a.copy()
    d = O.qr(f)
    return O.generate_q(f, d), O.qr_into_r(f, d)


def tsqr_chunks(a, ch):
    """[(Q_i, rows)] of the row chunks and the stacked R factors (the last chunk absorbs the remainder)."""
    m, n = a.shape
    nch = m // ch
    qs, rs = [], []
    for i in range(nch):
        r0, r1 = i * ch, (m if i == nch - 1 else (i + 1) * ch)
        q, r = local_q_r(a[r0:r1])
        qs.append(q)
        rs.append(r)
    return qs, np.vstack(rs)


def tsqr_explicit(a, ch):
    m, n = a.shape
    if m < 2 * ch:
        return local_q_r(a)
    qs, stack = tsqr_chunks(a, ch)
    Qs, R = tsqr_explicit(stack, ch)
    return np.vstack([q @ Qs[i * n:(i + 1) * n] for i, q in enumerate(qs)]), R


def lu_signs(q1):
    """In place LU of (Q1 - S): returns s; q1 then holds Y1 (strict lower, unit diagonal implied) and U (upper)."""
    n = q1.shape[0]
    s = np.zeros(n)
    for k in range(n):
        s[k] = -1.0 if q1[k, k] >= 0 else 1.0
        q1[k, k] -= s[k]
        q1[k + 1:, k] /= q1[k, k]
        q1[k + 1:, k + 1:] -= np.outer(q1[k + 1:, k], q1[k, k + 1:])
    return s


def finish_top(q1, s, R):
    """c, diag, U' = diag(1/c) U and the top block of the compact factor (csrc/tsqr_hr.cu: hr_scale / hr_finish)."""
    n = q1.shape[0]
    U = np.triu(q1)
    sp = np.concatenate(([1.0], s[:-1]))
    c = -sp * s * np.sqrt(np.abs(np.diag(U)) / 2)
    diag = sp * s * np.abs(np.diag(R))
    top = np.tril(q1, -1) * c[None, :] + np.diag(c) + np.triu(R, 1)
    return c, diag, U / c[:, None], top


def reconstruct(Q, R):
    n = Q.shape[1]
    q1 = Q[:n].copy()
    s = lu_signs(q1)
    c, diag, Up, top = finish_top(q1, s, R)
    rest = np.linalg.solve(Up.T, Q[n:].T).T            # Y2' = Q2 U'^-1
    return np.vstack([top, rest]), diag


def qr_tsqr_folded(a, ch):
    """only Q1 = Q_0[:n] Qs_0 is formed; the solve runs on the small stack Qs, and the per-chunk GEMM
    writes reflector rows directly."""
    m, n = a.shape
    qs, stack = tsqr_chunks(a, ch)
    Qs, R = tsqr_explicit(stack, ch)
    q1 = qs[0][:n] @ Qs[:n]
    s = lu_signs(q1)
    c, diag, Up, top = finish_top(q1, s, R)
    M = np.linalg.solve(Up.T, Qs.T).T                  # Qs U'^-1, (nch * n) x n
    out = np.vstack([q @ M[i * n:(i + 1) * n] for i, q in enumerate(qs)])
    out[:n] = top
    return out, diag


def check(a, f, d):
    ref = a.copy()
    dref = O.qr(ref)
    n = a.shape[1]
    eps = 2.220446049250313e-16
    nrm = np.linalg.norm(a, 2)
    assert np.max(np.abs(np.tril(f) - np.tril(ref))) <= 64 * n * eps
    assert np.max(np.abs(np.triu(f, 1) - np.triu(ref, 1))) <= 64 * n * eps * nrm
    assert np.max(np.abs(d - dref)) <= 64 * n * eps * nrm
    assert np.array_equal(np.signbit(d), np.signbit(dref))


@pytest.mark.parametrize("m,n,ch", [(400, 7, 50), (1000, 16, 64), (333, 5, 40), (130, 64, 1000), (5000, 24, 100)])
def test_reconstruction_reproduces_the_reference_factor(m, n, ch):
    a = np.random.default_rng(m + n).uniform(-100, 100, (m, n))
    Q, R = tsqr_explicit(a, ch)
    assert np.linalg.norm(Q.T @ Q - np.eye(n)) <= 64 * m * 2.3e-16
    f, d = reconstruct(Q, R)
    check(a, f, d)


@pytest.mark.parametrize("m,n,ch", [(400, 7, 50), (1000, 16, 64), (5000, 24, 100)])
def test_folded_solve_variant(m, n, ch):
    a = np.random.default_rng(7 * m + n).uniform(-100, 100, (m, n))
    f, d = qr_tsqr_folded(a, ch)
    check(a, f, d)


def test_zero_matrix_keeps_q_identity():
    """diag = +-0 carries the sign the reference's consumers read with signum (householder.rs:89): Q = [I; 0], R = 0."""
    a = np.zeros((200, 4))
    Q, R = tsqr_explicit(a, 30)
    f, d = reconstruct(Q, R)
    assert np.all(d == 0)
    assert np.array_equal(O.generate_q(f.copy(), d), np.eye(200, 4))
    assert np.all(O.qr_into_r(f.copy(), d) == 0)
