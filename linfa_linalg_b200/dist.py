"""Multi-GPU sharding of the two paths that shard naturally.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests):

* batched small-matrix QR: the batch index is split evenly, no collective;
* tall-skinny QR (TSQR): rank r owns a contiguous row block, factors it locally to an n x n R
  (diag >= 0, qr.rs:96), ONE all_gather of the R factors (n*n*8 bytes per rank, latency bound on
  NVSwitch), then every rank factors the stacked (world*n) x n matrix -> the global R, replicated.

* tall-skinny QR delivered as the reference's QRDecomp (`tsqr_qr`): the same row blocks; every rank
  keeps its explicit Q_r, the all-gathered R factors are reduced (replicated) to (Qs, R), Q_r <- Q_r Qs[r], the rank
  that owns the first n rows runs the Householder reconstruction of the top block (LU of Q - S) and broadcasts U' and
  diag (n*n + n values, one buffer), and every rank turns its rows into reflector rows with one right-hand TRSM.  Two
  collectives in all, both O(n^2) bytes.

Everything here is host-side plumbing; the arithmetic is behind the callables (`local_r`, `final_r`)
so that the same code path is exercised by the gloo tests with CPU stand-ins.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units owned by `rank`; the first total % world ranks get one more."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def stack_r_factors(r_all, world: int, n: int):
    """r_all: (world*n, n) row-major tensor whose block g is the ROW-MAJOR VIEW of rank g's column-major R_g
    (i.e. R_g^T).  Returns the (n, world*n) row-major tensor that IS the column-major stacked matrix
    [R_0; R_1; ...] ((world*n) x n, leading dimension world*n)."""
    return r_all.view(world, n, n).permute(1, 0, 2).reshape(n, world * n).contiguous()


def tsqr_r(local_r: Callable, final_r: Callable, n: int, group=None):
    """R factor of the row-sharded matrix.  `local_r()` -> (n, n) tensor holding this rank's column-major R
    (as torch sees it: R^T); `final_r(stack)` -> (n, n) tensor with the column-major R of the column-major
    stacked matrix `stack` ((n, world*n) row-major tensor).  With world == 1 no collective is issued."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    r = local_r()
    if world == 1:
        return r
    r_all = torch.empty((world * n, n), dtype=r.dtype, device=r.device)
    dist.all_gather_into_tensor(r_all, r.contiguous(), group=group)
    return final_r(stack_r_factors(r_all, world, n))


# ---- GPU callables over the C ABI ---------------------------------------------------------------------
def _bind_to_torch_stream(eng, device):
    """torch.distributed collectives (and torch's allocator) order against torch's CURRENT stream only, while an Engine
    runs on its own non-blocking stream by default.  Every GPU callable below therefore puts the engine on torch's
    current stream of `device` before it launches, so that kernels, `torch.empty` buffers and NCCL calls are all ordered
    on one stream (the switch drains the stream the engine was on before; a no-op when already bound)."""
    import torch
    eng.set_stream(torch.cuda.current_stream(device).cuda_stream)


def gpu_local_r(eng, block_cm, rows: int, n: int):
    """block_cm: torch f64 tensor of shape (n, rows) (row-major) == column-major rows x n block. Overwritten."""
    import torch
    r = torch.empty((n, n), dtype=torch.float64, device=block_cm.device)

    def run():
        _bind_to_torch_stream(eng, block_cm.device)
        st = eng.lib.lfb_tsqr_local_r_dev_f64(eng.h, C.c_void_p(block_cm.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n)
        if st != 0:
            raise RuntimeError(f"lfb_tsqr_local_r_dev_f64 status {st}: {eng.lib.lfb_last_error(eng.h)}")
        return r
    return run


def gpu_final_r(eng, n: int):
    import torch

    def run(stack):
        rows = stack.shape[1]
        _bind_to_torch_stream(eng, stack.device)
        r = torch.empty((n, n), dtype=torch.float64, device=stack.device)
        st = eng.lib.lfb_tsqr_local_r_dev_f64(eng.h, C.c_void_p(stack.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n)
        if st != 0:
            raise RuntimeError(f"lfb_tsqr_local_r_dev_f64 (stacked R) status {st}")
        return r
    return run


# ---- TSQR + Householder reconstruction over row blocks ---------------------------------------------------
def tsqr_qr(block, ops, n: int, group=None):
    """qr.rs:29-45 of a row-sharded tall-skinny matrix.  `block`: this rank's rows as an (n, rows_local) row-major
    tensor (== the column-major rows_local x n block); overwritten with ITS ROWS of the reference's compact factor
    (unit-norm reflectors below the diagonal, R above, householder.rs:34-51).  Returns (diag, r): the signed pivots
    (n) and the column-major R (diag >= 0) as an (n, n) tensor, both replicated.  `ops` supplies the arithmetic
    (GpuTsqrOps over the C ABI; CPU stand-ins in the gloo tests):
        explicit_q(x) -> r            x (n, rows) <- explicit thin Q of x, r = its column-major R
        apply_q(x, qs, row0=0)        rows row0.. of x <- (those rows) * qs     (qs: (n, n) tensor, column-major n x n)
        leaf(x) -> (r, rinv) | None   optional: Cholesky-QR leaf, x untouched; None when the block is declined
        matmul(a, b) -> c             optional (with leaf): column-major product of two small column-major operands
        reconstruct_top(x, r, u, diag)   first n rows of x -> top block of the compact factor; u <- U', diag
        reconstruct_rows(x, row0, u)  rows row0.. of x <- (those rows) U'^-1
    Every rank must own at least n rows (checked collectively: ValueError on all ranks otherwise).
    Stream contract: the collectives run on torch's current stream; `GpuTsqrOps` binds the engine to that stream on every
    call, so no extra synchronisation is needed between the engine's kernels and NCCL."""
    import torch
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    rows_local = block.shape[1]
    # every rank factors a thin block (qr.rs:34-36 NotThin per block) and rank 0 must hold the whole top n x n block:
    # agree on that BEFORE the first data collective, so a short shard raises on every rank instead of hanging the others
    ok = torch.tensor([1 if rows_local >= n else 0], dtype=torch.int32, device=block.device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        raise ValueError(f"tsqr_qr needs at least n = {n} rows on every rank (this rank owns {rows_local})")
    # The Cholesky-QR leaf (csrc/cholqr.cu) leaves the block untouched and returns R_i, R_i^-1: if EVERY rank's block is
    # accepted, everything after it folds into n x n products and the rows are touched by one more GEMM only:
    #   rows <- rows W_i,  W_i = R_i^-1 Qs_i U'^-1       (otherwise: explicit Q, Q_i <- Q_i Qs_i, right-hand TRSM)
    folded, rinv = False, None
    if hasattr(ops, "leaf"):
        got = ops.leaf(block)
        flag = torch.tensor([1 if got is not None else 0], dtype=torch.int32, device=block.device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        folded = int(flag.item()) == 1
        if folded:
            r, rinv = got
    if not folded:
        r = ops.explicit_q(block, householder_only=True) if hasattr(ops, "leaf") else ops.explicit_q(block)
    m_i = rinv                                                          # M_i = R_i^-1 Qs_i (world 1: Qs = I)
    if world > 1:
        r_all = torch.empty((world * n, n), dtype=r.dtype, device=r.device)
        dist.all_gather_into_tensor(r_all, r.contiguous(), group=group)
        stack = stack_r_factors(r_all, world, n)
        r = ops.explicit_q(stack)                                   # stack <- Qs, replicated
        qs_i = stack[:, rank * n:(rank + 1) * n].contiguous()
        if folded:
            m_i = ops.matmul(rinv, qs_i)
        else:
            ops.apply_q(block, qs_i)
    ud = torch.empty((n + 1, n), dtype=block.dtype, device=block.device)     # U' and diag share one buffer: ONE broadcast
    u, diag = ud[:n], ud[n]
    top = None
    if rank == 0:
        if folded:
            top = ops.matmul(block[:, :n], m_i)                     # Q_top = A_top M_0 (n x n), out of place
            ops.reconstruct_top(top, r, u, diag)
        else:
            ops.reconstruct_top(block, r, u, diag)
    if world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(ud, src=src, group=group)
    if folded:
        ops.reconstruct_rows(m_i, 0, u)                             # W_i = M_i U'^-1 (n x n)
        ops.apply_q(block, m_i, row0=n if rank == 0 else 0)         # rows <- rows W_i
        if rank == 0:
            block[:, :n].copy_(top)
    else:
        ops.reconstruct_rows(block, n if rank == 0 else 0, u)
    return diag, r


class GpuTsqrOps:
    """The four steps of `tsqr_qr` on liblinfa_b200 (f64, device tensors, the engine's stream)."""

    def __init__(self, eng):
        self.eng = eng

    def _call(self, name, *args):
        import torch
        _bind_to_torch_stream(self.eng, torch.device("cuda", self.eng.device))
        st = getattr(self.eng.lib, name)(self.eng.h, *args)
        if st != 0:
            raise RuntimeError(f"{name} status {st}: {self.eng.lib.lfb_last_error(self.eng.h)}")

    def explicit_q(self, x, householder_only=False):
        import torch
        n, rows = x.shape
        r = torch.empty((n, n), dtype=torch.float64, device=x.device)
        keep = None
        if householder_only:          # the leaf has just declined one of the blocks: do not repeat the attempt
            keep = 16
            self.eng.set_option("tsqr_cholqr_cond", 0)
        try:
            self._call("lfb_tsqr_explicit_q_dev_f64", C.c_void_p(x.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n)
        finally:
            if keep is not None:
                self.eng.set_option("tsqr_cholqr_cond", keep)
        return r

    def leaf(self, x):
        import torch
        n, rows = x.shape
        r = torch.empty((n, n), dtype=torch.float64, device=x.device)
        rinv = torch.empty((n, n), dtype=torch.float64, device=x.device)
        ok = C.c_int(0)
        self._call("lfb_tsqr_leaf_dev_f64", C.c_void_p(x.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n,
                   C.c_void_p(rinv.data_ptr()), n, C.byref(ok))
        return (r, rinv) if ok.value == 1 else None

    def matmul(self, a, b):
        """c = a b for column-major operands held as row-major tensors: a is (k, m) [column-major m x k, leading dimension
        a.stride(0)], b is (n, k) contiguous [column-major k x n]; returns the (n, m) tensor of the column-major m x n product."""
        import torch
        k, m = a.shape
        n = b.shape[0]
        c = torch.empty((n, m), dtype=torch.float64, device=b.device)
        self._call("lfb_gemm_dev_f64", 0, 0, m, n, k, 1.0, C.c_void_p(a.data_ptr()), a.stride(0), C.c_void_p(b.data_ptr()), b.stride(0),
                   0.0, C.c_void_p(c.data_ptr()), m)
        return c

    def apply_q(self, x, qs, row0=0):
        n, rows = x.shape
        if rows - row0 <= 0:
            return
        self._call("lfb_tsqr_apply_q_dev_f64", C.c_void_p(x.data_ptr() + row0 * x.element_size()), rows - row0, n, rows,
                   C.c_void_p(qs.data_ptr()), n)

    def reconstruct_top(self, x, r, u, diag):
        n, rows = x.shape
        self._call("lfb_hh_reconstruct_top_dev_f64", C.c_void_p(x.data_ptr()), n, rows, C.c_void_p(r.data_ptr()), n,
                   C.c_void_p(u.data_ptr()), n, C.c_void_p(diag.data_ptr()))

    def reconstruct_rows(self, x, row0, u):
        n, rows = x.shape
        if rows - row0 <= 0:
            return
        self._call("lfb_hh_reconstruct_rows_dev_f64", C.c_void_p(x.data_ptr() + row0 * x.element_size()), rows - row0, n, rows,
                   C.c_void_p(u.data_ptr()), n)


# ---- single-process multi-GPU over the C ABI (csrc/multi.cu) ------------------------------------------------
class MultiEngine:
    """One `lfb_multi` handle: several devices of one box driven from ONE process (one host thread + one engine handle
    per device inside the library, ncclCommInitAll for the R-factor exchange).  This is the launch mode a Rust caller of
    the shim has; the torch.distributed functions above are the one-process-per-GPU mode of bench.py and use the same
    split (`shard_range`).  The methods mirror linfa_linalg_b200's single-device functions."""

    def __init__(self, devices=None, n_devices=None):
        from . import _ffi
        self.lib = _ffi.load()
        if devices is None:
            if n_devices is None:
                raise ValueError("give a device list or n_devices")
            devices = list(range(n_devices))
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        hp = C.c_void_p()
        st = self.lib.lfb_create_multi(C.byref(hp), arr, len(self.devices))
        if st != 0:
            raise RuntimeError(f"lfb_create_multi({self.devices}) failed with status {st} (100 = CUDA, 102 = NCCL; no CPU fallback)")
        self.h = hp

    def close(self):
        if getattr(self, "h", None):
            self.lib.lfb_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise RuntimeError(f"liblinfa_b200 (multi) status {st}: {(self.lib.lfb_multi_last_error(self.h) or b'').decode()}")

    @property
    def nccl_ranks(self) -> int:
        return int(self.lib.lfb_multi_nccl_ranks(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.lfb_multi_launch_count(self.h))

    def set_option(self, key: str, value: int):
        self._check(self.lib.lfb_multi_set_option(self.h, key.encode(), int(value)))

    def synchronize(self):
        self._check(self.lib.lfb_multi_synchronize(self.h))

    @staticmethod
    def _sfx(a):
        import numpy as np
        if a.dtype == np.float64:
            return "_f64"
        if a.dtype == np.float32:
            return "_f32"
        raise TypeError(f"A: NdFloat means f32 or f64, got {a.dtype}")

    # -- host views (numpy arrays of any strides), results in place like the `*_into` trait methods --
    def qr_tsqr_into(self, a):
        """qr.rs:29-45 on a tall-skinny `a`, rows sharded over the devices: `a` becomes the reference's compact factor;
        returns the signed pivots `diag`."""
        import numpy as np
        rows, cols = a.shape
        diag = np.zeros(cols, dtype=a.dtype)
        it = a.itemsize
        st = getattr(self.lib, "lfb_qr_tsqr_multi" + self._sfx(a))(self.h, C.c_void_p(a.ctypes.data), rows, cols, a.strides[0] // it,
                                                                   a.strides[1] // it, C.c_void_p(diag.ctypes.data))
        self._check(st)
        return diag

    def tsqr_r(self, a):
        """R (cols x cols, upper, diag >= 0; qr.rs:91-98) of the row-sharded `a`; `a` is not modified."""
        import numpy as np
        rows, cols = a.shape
        r = np.zeros((cols, cols), dtype=a.dtype)
        it = a.itemsize
        st = getattr(self.lib, "lfb_tsqr_r_multi" + self._sfx(a))(self.h, C.c_void_p(a.ctypes.data), rows, cols, a.strides[0] // it,
                                                                  a.strides[1] // it, C.c_void_p(r.ctypes.data), cols, 1)
        self._check(st)
        return r

    def qr_batched(self, a):
        """qr.rs:32-44 over a C-contiguous [batch][m][n] array, batch split over the devices; returns diag [batch][n]."""
        import numpy as np
        assert a.ndim == 3 and a.flags.c_contiguous
        batch, m, n = a.shape
        diag = np.zeros((batch, n), dtype=a.dtype)
        st = getattr(self.lib, "lfb_qr_batched_multi" + self._sfx(a))(self.h, C.c_void_p(a.ctypes.data), batch, m, n, C.c_void_p(diag.ctypes.data))
        self._check(st)
        return diag

    def cholesky_batched(self, a, clean: bool = True):
        """cholesky.rs:51-83 over a C-contiguous [batch][n][n] array, batch split over the devices.  Returns
        (fail_matrix, fail_index): (-1, -1) on success, else the first failing matrix in batch order and its row."""
        assert a.ndim == 3 and a.flags.c_contiguous and a.shape[1] == a.shape[2]
        fm, fi = C.c_int64(-1), C.c_int64(-1)
        st = getattr(self.lib, "lfb_cholesky_batched_multi" + self._sfx(a))(self.h, C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1],
                                                                            int(clean), C.byref(fm), C.byref(fi))
        if st == 1:
            return fm.value, fi.value
        self._check(st)
        return -1, -1

    # -- device-resident blocks (one torch tensor per device, (n, rows_i) row-major == column-major rows_i x n) --
    def _ptrs(self, tensors):
        return (C.c_void_p * len(tensors))(*[C.c_void_p(t.data_ptr()) if t is not None else None for t in tensors])

    def qr_tsqr_dev(self, blocks, diags=None, rs=None):
        n = blocks[0].shape[0]
        rows = (C.c_int64 * len(blocks))(*[b.shape[1] for b in blocks])
        self._check(self.lib.lfb_qr_tsqr_multi_dev_f64(self.h, self._ptrs(blocks), rows, n, rows,
                                                       self._ptrs(diags) if diags else None, self._ptrs(rs) if rs else None))

    def tsqr_r_dev(self, blocks, rs):
        n = blocks[0].shape[0]
        rows = (C.c_int64 * len(blocks))(*[b.shape[1] for b in blocks])
        self._check(self.lib.lfb_tsqr_r_multi_dev_f64(self.h, self._ptrs(blocks), rows, n, rows, self._ptrs(rs)))

    def qr_batched_dev_f32(self, mats, diags):
        m, n = mats[0].shape[1], mats[0].shape[2]
        batch = (C.c_int64 * len(mats))(*[t.shape[0] for t in mats])
        self._check(self.lib.lfb_qr_batched_multi_dev_f32(self.h, self._ptrs(mats), batch, m, n, self._ptrs(diags)))

    def time_begin(self):
        self._check(self.lib.lfb_multi_time_begin(self.h))

    def time_end(self) -> float:
        ms = C.c_double(0)
        self._check(self.lib.lfb_multi_time_end(self.h, C.byref(ms)))
        return ms.value
