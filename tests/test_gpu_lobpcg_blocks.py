"""GPU parity of the fused LOBPCG blocks (csrc/lobpcg_blocks.cu) against the oracle's
restatement of lobpcg/algorithm.rs:63-97.

Tolerances: u = v L^-T amplifies rounding by cond(v)^2 through the Gram matrix (the reference has the same property:
its own test uses 1e-2, algorithm.rs:494); on the well-conditioned random blocks used here
|u - u_oracle| <= 64 sqrt(rows) eps cond(v)^2 max|u| and |L - L_oracle| <= 16 sqrt(rows) eps cond(v) ||v||_2 (L is
O(||v||_2)); orthogonality and u L^T = v as the reference's own test checks them (algorithm.rs:492-502)."""
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


def rnd(shape, dt=np.float64, seed=0, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(dt)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("rows,cols", [(10, 10), (1000, 5), (5000, 64), (20000, 100), (3001, 130), (7, 1)])
def test_orthonormalize_parity(L, rows, cols, dt):
    v0 = rnd((rows, cols), dt, seed=rows + cols)
    st, fi, u_ref, l_ref = O.lobpcg_orthonormalize(v0.copy())
    assert st == 0
    for v in (v0.copy(), np.array(v0, order="F")):
        u, l = L.orthonormalize(v)
        eps = EPS[dt]
        nrm = np.linalg.norm(v0.astype(np.float64), 2)
        cond2 = (nrm / np.linalg.svd(v0.astype(np.float64), compute_uv=False)[-1]) ** 2
        assert np.all(np.triu(l, 1) == 0) and np.all(np.diag(l) > 0)
        assert np.max(np.abs(l - l_ref)) <= 16 * eps * np.sqrt(rows) * nrm * max(cond2, 1.0) ** 0.5
        assert np.max(np.abs(u - u_ref)) <= 64 * eps * np.sqrt(rows) * cond2 * np.max(np.abs(u_ref))
        # the reference's own properties (algorithm.rs:492-502): orthonormal columns, u L^T = v
        assert np.linalg.norm(u.T.astype(np.float64) @ u - np.eye(cols)) <= 64 * eps * rows ** 0.5 * cond2
        assert np.linalg.norm(u.astype(np.float64) @ l.T - v0) <= 64 * eps * rows ** 0.5 * np.linalg.norm(v0)


def test_orthonormalize_not_positive_definite(L):
    v = rnd((50, 8), seed=3)
    v[:, 5] = 0.0
    st, fi, _, _ = O.lobpcg_orthonormalize(v.copy())
    assert st == 1 and fi == 5
    with pytest.raises(L.NotPositiveDefinite) as ei:
        L.orthonormalize(v.copy())
    assert ei.value.index == 5
    u, l = L.orthonormalize(np.zeros((5, 0)))
    assert u.shape == (5, 0) and l.shape == (0, 0)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,k,m", [(30, 5, 4), (5000, 16, 3), (20000, 70, 130), (64, 64, 1)])
def test_apply_constraints_parity(L, n, k, m, dt):
    v0 = rnd((n, k), dt, seed=n + k)
    y = rnd((n, m), dt, seed=n + m + 1)
    lyy = np.ascontiguousarray((y.T @ y).astype(dt))
    assert O.cholesky(lyy)[0] == 0
    ref = O.lobpcg_apply_constraints(v0.copy(), lyy.copy(), y)
    for v in (v0.copy(), np.array(v0, order="F")):
        out = L.apply_constraints(v, lyy, y)
        assert out is v
        scale = np.max(np.abs(ref)) + np.linalg.norm(y.astype(np.float64), 2) * np.max(np.abs(np.linalg.solve(np.tril(lyy).astype(np.float64), (y.T @ v0).astype(np.float64))))
        assert np.max(np.abs(out - ref)) <= 64 * EPS[dt] * np.sqrt(n) * scale
    with pytest.raises(ValueError):
        L.apply_constraints(v0.copy(), lyy, y[:-1])


def test_lobpcg_blocks_device_resident(L):
    """The device entry points chained without leaving HBM: constraints, then orthonormalisation (algorithm.rs:157-166)."""
    import torch
    e = L.Engine(0)
    n, k, m = 40000, 32, 8
    v0 = rnd((n, k), seed=1)
    y = np.linalg.qr(rnd((n, m), seed=2))[0]
    lyy = np.ascontiguousarray(y.T @ y)
    assert O.cholesky(lyy)[0] == 0
    ref = O.lobpcg_apply_constraints(v0.copy(), lyy.copy(), y)
    st, _, u_ref, l_ref = O.lobpcg_orthonormalize(ref)
    assert st == 0
    V = torch.from_numpy(np.ascontiguousarray(v0.T)).cuda()           # (k, n) row-major == column-major n x k
    Y = torch.from_numpy(np.ascontiguousarray(y.T)).cuda()
    Ly = torch.from_numpy(np.ascontiguousarray(lyy.T)).cuda()
    Lm = torch.zeros((k, k), dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e._check(e.call("lfb_apply_constraints_dev_f64", C.c_void_p(V.data_ptr()), n, k, n, C.c_void_p(Ly.data_ptr()), m, m,
                    C.c_void_p(Y.data_ptr()), n))
    e._check(e.call("lfb_orthonormalize_dev_f64", C.c_void_p(V.data_ptr()), n, k, n, C.c_void_p(Lm.data_ptr()), k,
                    C.c_void_p(info.data_ptr())))
    torch.cuda.synchronize()
    e.set_stream(None)
    assert int(info[0]) == 0
    u = V.t().cpu().numpy()
    l = Lm.t().cpu().numpy()
    assert np.max(np.abs(u - u_ref)) <= 1e-12 and np.max(np.abs(l - l_ref)) <= 1e-10
    assert np.max(np.abs(y.T @ u)) <= 1e-12 and np.linalg.norm(u.T @ u - np.eye(k)) <= 1e-12
    e.close()
