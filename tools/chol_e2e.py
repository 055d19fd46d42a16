"""Host-view Cholesky end to end (H2D + factorisation + D2H inside the timed region) for several `chol_waves` settings in ONE
process (box-to-box PCIe / host variance is larger than the effect).  usage: python tools/chol_e2e.py [n] [reps] [pinned|pageable] [C|F]"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "pinned"
order = sys.argv[4] if len(sys.argv) > 4 else "C"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
S = (S + S.t()) / 2
S.diagonal().add_(float(n))
src = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
src.copy_(S)
work = torch.empty((n, n), dtype=torch.float64, pin_memory=(kind == "pinned"))
rs, cs = (n, 1) if order == "C" else (1, n)          # the matrix is symmetric: the same buffer read either way
fail = C.c_int64(-1)
for waves in (1, 3, 1, 3, 2, 4):
    eng = L.Engine(0)
    eng.set_option("chol_waves", waves)
    best, tot = 1e30, 0.0
    for it in range(reps + 1):
        work.copy_(src)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.lib.lfb_cholesky_f64(eng.h, C.c_void_p(work.data_ptr()), n, n, rs, cs, 0, C.byref(fail))
        dt = (time.perf_counter() - t0) * 1e3
        assert st == 0, st
        if it > 0:
            best = min(best, dt)
            tot += dt
    k = min(n, 2048)
    blk = work[:k, :k] if order == "C" else work[:k, :k].t()
    Lh = torch.tril(blk).to(dev)
    Lt = torch.tril(work[-k:, -k:] if order == "C" else work[-k:, -k:].t()).to(dev)      # trailing block: every wave has touched it
    res = float((Lh @ Lh.t() - S[:k, :k]).norm() / S[:k, :k].norm())
    print(json.dumps({"n": n, "host": kind, "order": order, "chol_waves": waves, "best_ms": round(best, 2), "mean_ms": round(tot / reps, 2),
                      "gflops_best": round(n ** 3 / 3 / best / 1e6, 1), "resid_leading_block": res, "trailing_diag_min": float(Lt.diagonal().min())}), flush=True)
