"""a LOBPCG iteration (src/lobpcg/algorithm.rs:119-431) with EVERY n x k block resident in HBM.

The loop below is the reference's loop (dense operator, no preconditioner, no constraints) written over device tensors and
the `*_dev` entry points of the C ABI only: `lfb_orthonormalize_dev_f64` (:81-97), `lfb_sorted_eig_dev_f64` (:16-44),
`lfb_hh_reconstruct_rows_dev_f64` (the right-hand triangular solve of :314-320) and `lfb_gemm_dev_f64` for every block
product.  Host traffic per iteration: k residual norms and k eigenvalues (the control decisions the reference also makes
on scalars) -- no block ever crosses PCIe.  Checked against tests/lobpcg_ref.py (the same sequence in NumPy): the
eigenvalue trajectory iteration by iteration, and the converged eigenpairs against LAPACK."""
import ctypes as C

import numpy as np
import pytest

from lobpcg_ref import lobpcg as lobpcg_numpy

pytestmark = pytest.mark.gpu


class Dev:
    """Column-major device matrices as row-major torch tensors (tensor of shape (cols, rows))."""

    def __init__(self, L):
        import torch
        self.torch = torch
        self.e = L.Engine(0)
        self.e.set_stream(torch.cuda.current_stream().cuda_stream)
        self.lib = self.e.lib
        self.info = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.block_h2d = 0          # bytes of n x k blocks uploaded after the start (must stay 0)

    def p(self, t):
        return C.c_void_p(t.data_ptr())

    def mm(self, a, b, ta=0, tb=0, alpha=1.0, beta=0.0, out=None):
        """column-major  out = alpha op(a) op(b) + beta out  (operands: tensors (cols, rows) contiguous)"""
        torch = self.torch
        m = a.shape[0] if ta else a.shape[1]
        k = a.shape[1] if ta else a.shape[0]
        n = b.shape[1] if tb else b.shape[0]
        if out is None:
            out = torch.empty((n, m), dtype=torch.float64, device="cuda")
        st = self.lib.lfb_gemm_dev_f64(self.e.h, ta, tb, m, n, k, alpha, self.p(a), a.shape[1], self.p(b), b.shape[1], beta,
                                       self.p(out), m)
        assert st == 0
        return out

    def orthonormalize(self, v):
        k, n = v.shape
        l = self.torch.empty((k, k), dtype=self.torch.float64, device="cuda")
        st = self.lib.lfb_orthonormalize_dev_f64(self.e.h, self.p(v), n, k, n, self.p(l), k, self.p(self.info))
        assert st == 0
        if int(self.info.item()) != 0:
            raise np.linalg.LinAlgError("NotPositiveDefinite")
        return v, l            # v overwritten with u; l: column-major L

    def right_solve_lt(self, b, l):
        """b <- b L^-T  (algorithm.rs:314-320: p_r.solve_triangular_into(ap^T, Lower)^T)"""
        k, n = b.shape
        u = l.t().contiguous()             # column-major U = L^T
        st = self.lib.lfb_hh_reconstruct_rows_dev_f64(self.e.h, self.p(b), n, k, n, self.p(u), k)
        assert st == 0
        return b

    def sorted_eig(self, ga, gb, size, largest=True):
        torch = self.torch
        k = ga.shape[0]
        vals = np.zeros(size)
        vecs = torch.empty((size, k), dtype=torch.float64, device="cuda")        # column-major k x size
        st = self.lib.lfb_sorted_eig_dev_f64(self.e.h, self.p(ga), k, self.p(gb) if gb is not None else None, k, k, size,
                                             1 if largest else 2, C.c_void_p(vals.ctypes.data), self.p(vecs), k)
        assert st == 0
        return vals, vecs


def lobpcg_resident(D, A, x0, tol, maxiter):
    torch = D.torch
    n, size_x = x0.shape
    it = min(n * 10, maxiter)
    X = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()                 # the ONLY block upload
    X, _ = D.orthonormalize(X)
    AX = D.mm(A, X)
    lam, eb = D.sorted_eig(D.mm(X, AX, ta=1), None, size_x)
    X, AX = D.mm(X, eb), D.mm(AX, eb)
    active = np.ones(size_x, dtype=bool)
    prev = None
    hist = [lam.copy()]
    blk = lambda rows: torch.cat([torch.cat(r, dim=1) for r in rows], dim=0).contiguous()   # block rows; the matrices are symmetric
    sym = lambda g: (g + g.t()) / 2
    while True:
        lam_d = torch.from_numpy(lam).cuda()
        R = AX - X * lam_d[:, None]                                          # column j scaled by lambda_j
        rn = R.norm(dim=1).cpu().numpy()                                     # k scalars to the host: the stopping test
        active = (rn > tol) & active
        cur = int(active.sum())
        if cur == 0 or it == 0:
            break
        idx = torch.from_numpy(np.nonzero(active)[0]).cuda()
        Ra = R[idx].contiguous()
        D.mm(X, D.mm(X, Ra, ta=1), alpha=-1.0, beta=1.0, out=Ra)             # Ra -= X (X^T Ra)
        Rn, _ = D.orthonormalize(Ra)
        AR = D.mm(A, Rn)
        xar, rar = D.mm(X, AR, ta=1), sym(D.mm(Rn, AR, ta=1))
        xax, xx, rr, xr = sym(D.mm(X, AX, ta=1)), D.mm(X, X, ta=1), D.mm(Rn, Rn, ta=1), D.mm(X, Rn, ta=1)
        p_ap = None
        if prev is not None:
            P, AP = prev
            try:
                Pa, l = D.orthonormalize(P[idx].contiguous())
                p_ap = (Pa, D.right_solve_lt(AP[idx].contiguous(), l))
            except np.linalg.LinAlgError:
                p_ap = None
        T = lambda g: g.t()
        if p_ap is not None:
            Pa, APa = p_ap
            xap, rap, pap = D.mm(X, APa, ta=1), D.mm(Rn, APa, ta=1), sym(D.mm(Pa, APa, ta=1))
            xp, rp, pp = D.mm(X, Pa, ta=1), D.mm(Rn, Pa, ta=1), D.mm(Pa, Pa, ta=1)
            # tensors hold the TRANSPOSE of each column-major block: the block matrix is assembled on its transpose
            ga = blk([[xax, T(xar), T(xap)], [xar, rar, T(rap)], [xap, rap, pap]])
            gb = blk([[xx, T(xr), T(xp)], [xr, rr, T(rp)], [xp, rp, pp]])
        else:
            ga = blk([[xax, T(xar)], [xar, rar]])
            gb = blk([[xx, T(xr)], [xr, rr]])
        lam, ev = D.sorted_eig(ga, gb, size_x)                               # ev: tensor (size_x, K) = column-major K x size_x
        tau = ev[:, :size_x].contiguous()
        if p_ap is not None:
            Pa, APa = p_ap
            alpha, gamma = ev[:, size_x:size_x + cur].contiguous(), ev[:, size_x + cur:].contiguous()
            P = D.mm(Pa, gamma, out=D.mm(Rn, alpha), beta=1.0)
            AP = D.mm(APa, gamma, out=D.mm(AR, alpha), beta=1.0)
        else:
            alpha = ev[:, size_x:].contiguous()
            P, AP = D.mm(Rn, alpha), D.mm(AR, alpha)
        X = D.mm(X, tau) + P
        AX = D.mm(AX, tau) + AP
        prev = (P, AP)
        hist.append(lam.copy())
        it -= 1
    return hist, lam, X, rn


@pytest.mark.parametrize("n,k", [(600, 4), (2000, 8)])
def test_lobpcg_loop_with_resident_blocks(n, k):
    import torch
    import linfa_linalg_b200 as L
    rng = np.random.default_rng(n + k)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    spec = np.concatenate([np.linspace(1.0, 2.0, n - k), 10.0 + 3.0 * np.arange(1, k + 1)])      # k well separated top eigenvalues
    a = (q * spec) @ q.T
    a = (a + a.T) / 2
    x0 = rng.standard_normal((n, k))
    tol, maxiter = 1e-8, 40
    hist_ref, lam_ref, x_ref, rhist_ref = lobpcg_numpy(a, x0.copy(), tol, maxiter)
    D = Dev(L)
    A = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    l0 = D.e.launch_count
    hist, lam, X, rn = lobpcg_resident(D, A, x0, tol, maxiter)
    torch.cuda.synchronize()
    assert D.e.launch_count > l0
    top = np.sort(spec)[::-1][:k]
    assert np.max(np.abs(lam - top)) <= 1e-8 * top[0]                        # converged to the k largest eigenvalues
    assert np.max(rn) <= 1e-6
    xs = X.cpu().numpy().T
    assert np.linalg.norm(a @ xs - xs * lam[None, :]) <= 1e-6 * top[0]
    # the trajectory follows the reference's loop: same eigenvalue estimates iteration by iteration while the run is far
    # from convergence (afterwards both sit at the answer and differ by rounding-level noise over noise)
    m = min(len(hist), len(hist_ref), 6)
    for i in range(m):
        assert np.max(np.abs(hist[i] - hist_ref[i])) <= 1e-7 * top[0], (i, hist[i], hist_ref[i])
    assert abs(len(hist) - len(hist_ref)) <= 2
    D.e.set_stream(None)
    D.e.close()


def test_sorted_eig_dev_and_host_agree_with_numpy():
    """lfb_sorted_eig_* (host views) and lfb_sorted_eig_dev_f64 on a generalized problem: algorithm.rs:505-522 style."""
    import torch
    import linfa_linalg_b200 as L
    rng = np.random.default_rng(9)
    k = 48
    g = rng.standard_normal((k, k)); a = (g + g.T) / 2
    h = rng.standard_normal((k, k)); b = h @ h.T + k * np.eye(k)
    import scipy.linalg as sl
    ref = sl.eigh(a, b, eigvals_only=True)[::-1]
    vals, vecs = L.sorted_eig(a.copy(), b.copy(), 10, L.LARGEST)
    assert vals.shape == (10,) and vecs.shape == (k, 10)
    assert np.max(np.abs(vals - ref[:10])) <= 1e-10 * np.abs(ref).max()
    assert not np.any(np.signbit(vecs[0, :]))
    assert np.linalg.norm(a @ vecs - b @ vecs * vals[None, :]) <= 1e-9 * np.linalg.norm(a)
    v0, q0 = L.generalized_eig(a.copy(), b.copy())
    assert np.max(np.abs(np.sort(v0)[::-1] - ref)) <= 1e-10 * np.abs(ref).max()
    D = Dev(L)
    va, qa = D.sorted_eig(torch.from_numpy(a.T.copy()).cuda(), torch.from_numpy(b.T.copy()).cuda(), 10, True)
    torch.cuda.synchronize()
    assert np.max(np.abs(va - vals)) <= 1e-12 * np.abs(ref).max()
    assert np.max(np.abs(qa.cpu().numpy().T - vecs)) <= 1e-10
    for dt in (np.float32,):
        v32, q32 = L.sorted_eig(a.astype(dt), None, 5, L.SMALLEST)
        assert np.max(np.abs(v32 - np.linalg.eigvalsh(a)[:5])) <= 64 * k * 1.2e-7 * np.linalg.norm(a, 2)
    D.e.set_stream(None)
    D.e.close()
