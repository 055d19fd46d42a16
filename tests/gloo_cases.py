"""(Run through tests/test_dist_gloo.py, one fresh interpreter per case.)  CPU, world_size 2, gloo: the host-side sharding logic of the two sharded paths.
The local arithmetic is the oracle here (tests may use it); on the GPU box the same dist.py code path
runs with the CUDA callables (bench.py extras)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linfa_linalg_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 262144, 4194304):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _oracle_r_cm(a):
    """column-major R (as a row-major torch tensor: R^T) of the numpy matrix a"""
    import oracle as O
    w = np.array(a, dtype=np.float64, order="C")
    d = O.qr(w)
    return torch.from_numpy(np.ascontiguousarray(O.qr_into_r(w, d).T))


def _tsqr_worker(rank, world, port, rows, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.random.default_rng(7).uniform(-1, 1, (rows, n))
    b, e = D.shard_range(rows, world, rank)
    local = full[b:e]

    def final_r(stack):  # stack: (n, world*n) row-major == column-major (world*n) x n
        return _oracle_r_cm(stack.numpy().T)

    r = D.tsqr_r(lambda: _oracle_r_cm(local), final_r, n)
    out[rank] = r.numpy().T.copy()
    # batched path: shards are disjoint and cover the batch
    bb, be = D.shard_range(1000, world, rank)
    cnt = torch.tensor([be - bb])
    dist.all_reduce(cnt)
    assert int(cnt) == 1000
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_tsqr_two_ranks_matches_single_qr():
    import oracle as O
    rows, n, world = 400, 12, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_tsqr_worker, args=(world, _free_port(), rows, n, out), nprocs=world, join=True)
    full = np.random.default_rng(7).uniform(-1, 1, (rows, n))
    w = full.copy()
    d = O.qr(w)
    r_ref = O.qr_into_r(w, d)
    for rank in range(world):
        r = out[rank]
        assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
        assert np.max(np.abs(r - r_ref)) <= 64 * n * 2.2e-16 * np.linalg.norm(full, 2)


def test_stack_layout():
    world, n = 3, 4
    rs = [torch.arange(n * n, dtype=torch.float64).reshape(n, n) + 100 * g for g in range(world)]   # row-major views R_g^T
    stack = D.stack_r_factors(torch.cat(rs, 0), world, n)
    # column-major stacked matrix X = [R_0; R_1; R_2]: X[g*n + r, c] == R_g[r, c] == rs[g][c, r]
    X = stack.t()
    for g in range(world):
        assert torch.equal(X[g * n:(g + 1) * n, :], rs[g].t())


# ---- tsqr_qr: TSQR + Householder reconstruction over row blocks (dist.py), CPU stand-ins for the four device steps ----
class _CpuTsqrOps:
    """NumPy/oracle restatement of csrc/tsqr_hr.cu's steps on the same tensor layout ((n, rows) row-major tensors ==
    column-major rows x n blocks), so the gloo test exercises exactly the plumbing the GPU path uses."""

    @staticmethod
    def _cm(x):  # numpy view of the column-major matrix (rows x n), writable, shares memory with the tensor
        return x.numpy().T

    def explicit_q(self, x):
        import oracle as O
        a = self._cm(x)
        f = np.array(a, dtype=np.float64, order="C")
        d = O.qr(f)
        a[:] = O.generate_q(f, d)
        return torch.from_numpy(np.ascontiguousarray(O.qr_into_r(f, d).T))

    def apply_q(self, x, qs):
        a = self._cm(x)
        a[:] = a @ qs.numpy().T

    def reconstruct_top(self, x, r, u, diag):
        n = x.shape[0]
        q = self._cm(x)[:n]          # top n x n block (view)
        R = r.numpy().T
        s = np.zeros(n)
        for k in range(n):
            s[k] = -1.0 if q[k, k] >= 0 else 1.0
            q[k, k] -= s[k]
            q[k + 1:, k] /= q[k, k]
            q[k + 1:, k + 1:] -= np.outer(q[k + 1:, k], q[k, k + 1:])
        U = np.triu(q).copy()
        sp = 1.0
        c = np.zeros(n)
        for k in range(n):
            c[k] = -sp * s[k] * np.sqrt(abs(U[k, k]) / 2)
            diag[k] = sp * s[k] * R[k, k]
            sp = s[k]
        for k in range(n):
            q[k + 1:, k] *= c[k]
            q[k, k] = c[k]
            q[k, k + 1:] = R[k, k + 1:]
        u.copy_(torch.from_numpy(np.ascontiguousarray((U / c[:, None]).T)))

    def reconstruct_rows(self, x, row0, u):
        a = self._cm(x)
        if a.shape[0] - row0 <= 0:
            return
        U = u.numpy().T
        a[row0:] = np.linalg.solve(U.T, a[row0:].T).T


class _CpuFoldedOps(_CpuTsqrOps):
    """The same stand-ins plus the optional Cholesky-QR leaf (csrc/cholqr.cu) and the small column-major product, so that
    the FOLDED route of dist.tsqr_qr (rows <- rows R_i^-1 Qs_i U'^-1) runs under gloo too.  `decline_on_rank` makes one
    rank's leaf refuse its block: every rank must then fall back together."""

    def __init__(self, decline_on_rank=None):
        self.decline_on_rank = decline_on_rank
        self.leaf_calls = 0

    def explicit_q(self, x, householder_only=False):
        return super().explicit_q(x)

    def leaf(self, x):
        self.leaf_calls += 1
        if self.decline_on_rank is not None and dist.is_initialized() and dist.get_rank() == self.decline_on_rank:
            return None
        a = self._cm(x)
        l = np.linalg.cholesky(a.T @ a)
        r = l.T
        return torch.from_numpy(np.ascontiguousarray(r.T)), torch.from_numpy(np.ascontiguousarray(np.linalg.inv(r).T))

    def matmul(self, a, b):
        return torch.from_numpy(np.ascontiguousarray((a.numpy().T @ b.numpy().T).T))

    def apply_q(self, x, qs, row0=0):
        a = self._cm(x)
        a[row0:] = a[row0:] @ qs.numpy().T


def _make_ops(kind):
    return {"plain": _CpuTsqrOps, "folded": _CpuFoldedOps, "declined": lambda: _CpuFoldedOps(decline_on_rank=1)}[kind]()


def _tsqr_qr_worker(rank, world, port, rows, n, out, kind="plain"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.random.default_rng(11).uniform(-100, 100, (rows, n))
    b, e = D.shard_range(rows, world, rank)
    block = torch.from_numpy(np.ascontiguousarray(full[b:e].T))     # (n, rows_local) row-major == column-major block
    diag, r = D.tsqr_qr(block, _make_ops(kind), n)
    out[rank] = (block.numpy().T.copy(), diag.numpy().copy(), r.numpy().T.copy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind", ["plain", "folded", "declined"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_tsqr_qr_ranks_match_reference_compact_factor(world, kind):
    """The row blocks returned by tsqr_qr, stacked, are the reference's compact QR factor of the whole matrix
    (qr.rs:29-45 via the oracle): reflectors, R rows and signed pivots, elementwise."""
    import oracle as O
    rows, n = 301, 9
    full = np.random.default_rng(11).uniform(-100, 100, (rows, n))
    if world == 1:
        block = torch.from_numpy(np.ascontiguousarray(full.T))
        diag, r = D.tsqr_qr(block, _make_ops(kind), n)
        got = {0: (block.numpy().T.copy(), diag.numpy().copy(), r.numpy().T.copy())}
    else:
        mgr = mp.Manager()
        got = mgr.dict()
        mp.spawn(_tsqr_qr_worker, args=(world, _free_port(), rows, n, got, kind), nprocs=world, join=True)
    ref = full.copy()
    dref = O.qr(ref)
    tol = 64 * rows * 2.2e-16 * np.linalg.norm(full, 2)
    factor = np.vstack([got[rk][0] for rk in range(world)])
    assert factor.shape == ref.shape
    assert np.max(np.abs(factor - ref)) <= tol
    for rk in range(world):
        assert np.max(np.abs(got[rk][1] - dref)) <= tol
        assert np.max(np.abs(got[rk][2] - O.qr_into_r(ref.copy(), dref))) <= tol
    # and the reference's consumers accept it: Q from the stacked factor reproduces A (tests/qr.rs:20-27)
    q = O.generate_q(factor.copy(), got[0][1])
    rr = O.qr_into_r(factor.copy(), got[0][1])
    assert np.linalg.norm(q @ rr - full) <= tol * n
    assert np.linalg.norm(q.T @ q - np.eye(n)) <= 64 * rows * 2.2e-16


def _tsqr_qr_short_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    rows = 40 if rank == 0 else 3                       # rank 1's shard is shorter than n
    block = torch.from_numpy(np.random.default_rng(rank).uniform(-1, 1, (n, rows)))
    try:
        D.tsqr_qr(block, _CpuTsqrOps(), n)
        out[rank] = "no error"
    except ValueError as ex:
        out[rank] = str(ex)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_tsqr_qr_short_shard_raises_on_every_rank():
    """A shard with fewer rows than columns cannot be factored thin (qr.rs:34-36): both ranks raise, nobody hangs."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_tsqr_qr_short_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert "at least n = 6 rows" in out[0] and "at least n = 6 rows" in out[1]
