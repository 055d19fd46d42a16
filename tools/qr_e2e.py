"""Host-view QR end to end (H2D + factorisation + D2H inside the timed region) with and without the overlapped download, in ONE
process.  usage: python tools/qr_e2e.py [n] [reps]"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
src = torch.rand((n, n), dtype=torch.float64).mul_(2).sub_(1).pin_memory()
work = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
diag = torch.empty(n, dtype=torch.float64)
outs = {}
for ov in (0, 1, 0, 1):
    eng = L.Engine(0)
    eng.set_option("qr_overlap_d2h", ov)
    best = 1e30
    for it in range(reps + 1):
        work.copy_(src)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.lib.lfb_qr_f64(eng.h, C.c_void_p(work.data_ptr()), n, n, n, 1, C.c_void_p(diag.data_ptr()))
        dt = (time.perf_counter() - t0) * 1e3
        assert st == 0, st
        if it > 0:
            best = min(best, dt)
    outs[ov] = (work.clone(), diag.clone())
    print(json.dumps({"n": n, "qr_overlap_d2h": ov, "best_ms": round(best, 2), "gflops": round(4 / 3 * n ** 3 / best / 1e6, 1)}), flush=True)
print(json.dumps({"identical": bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))}))
