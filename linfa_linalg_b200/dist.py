"""Multi-GPU sharding of the two paths that shard naturally.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests):

* batched small-matrix QR: the batch index is split evenly, no collective;
* tall-skinny QR (TSQR): rank r owns a contiguous row block, factors it locally to an n x n R
  (diag >= 0, qr.rs:96), ONE all_gather of the R factors (n*n*8 bytes per rank, latency bound on
  NVSwitch), then every rank factors the stacked (world*n) x n matrix -> the global R, replicated.

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
def gpu_local_r(eng, block_cm, rows: int, n: int):
    """block_cm: torch f64 tensor of shape (n, rows) (row-major) == column-major rows x n block. Overwritten."""
    import torch
    r = torch.empty((n, n), dtype=torch.float64, device=block_cm.device)

    def run():
        st = eng.lib.lfb_tsqr_local_r_dev_f64(eng.h, C.c_void_p(block_cm.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n)
        if st != 0:
            raise RuntimeError(f"lfb_tsqr_local_r_dev_f64 status {st}: {eng.lib.lfb_last_error(eng.h)}")
        return r
    return run


def gpu_final_r(eng, n: int):
    import torch

    def run(stack):
        rows = stack.shape[1]
        r = torch.empty((n, n), dtype=torch.float64, device=stack.device)
        st = eng.lib.lfb_tsqr_local_r_dev_f64(eng.h, C.c_void_p(stack.data_ptr()), rows, n, rows, C.c_void_p(r.data_ptr()), n)
        if st != 0:
            raise RuntimeError(f"lfb_tsqr_local_r_dev_f64 (stacked R) status {st}")
        return r
    return run
