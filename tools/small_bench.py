"""Latency of the small single-CTA building blocks (potf2 / trsv) through the device API."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

eng = L.Engine(0)
lib = eng.lib
eng.set_stream(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda")
p = lambda t: C.c_void_p(t.data_ptr())
info = torch.zeros(1, dtype=torch.int64, device=dev)
for n in (64, 128, 256, 512, 1024):
    S = torch.rand((n, n), dtype=torch.float64, device=dev) * 2 - 1
    S = (S + S.t()) / 2
    S.diagonal().add_(float(n))
    reps = 50
    Ws = [S.clone() for _ in range(reps + 5)]
    for i in range(5):
        lib.lfb_cholesky_dev_f64(eng.h, p(Ws[i]), n, n, 0, p(info))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count
    e0.record()
    for i in range(reps):
        lib.lfb_cholesky_dev_f64(eng.h, p(Ws[5 + i]), n, n, 0, p(info))
    e1.record()
    torch.cuda.synchronize()
    print(f"cholesky n={n}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call, {(eng.launch_count - l0) // reps} launches")
