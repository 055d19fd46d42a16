"""Batched 32x32 f32 QR on the device: each `batched_quad` generation timed alone (CUDA events, min of reps, input restored
by a device copy outside the timed region) and held to generation 2's output and to the oracle on the first 512 matrices.
usage: python tools/batched_bench.py [batch] [reps] [quad,quad,...]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402
import oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
quads = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "2,4").split(",")]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(3)
S = torch.rand((B, 32, 32), dtype=torch.float32, device=dev, generator=g) * 2 - 1
S[1, :, 5] = 0.0                       # a `None` pivot
S[2] = 0.0
W = torch.empty_like(S)
D = torch.empty((B, 32), dtype=torch.float32, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
ref = S[:512].cpu().numpy().copy()
dref = O.qr_batched(ref)
base = None
for q in quads:
    eng = L.Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_option("batched_quad", q)
    best = 1e30
    for _ in range(reps + 3):
        W.copy_(S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = eng.lib.lfb_qr_batched_dev_f32(eng.h, p(W), B, 32, 32, p(D))
        e1.record()
        torch.cuda.synchronize()
        assert st == 0
        best = min(best, e0.elapsed_time(e1))
    err_o = float(max(np.max(np.abs(W[:512].cpu().numpy() - ref)), np.max(np.abs(D[:512].cpu().numpy() - dref))))
    res = {"batched_quad": q, "batch": B, "ms": round(best, 4), "matrices_per_s": B / best * 1e3,
           "GBps": B * 8320 / best / 1e6, "max_abs_vs_oracle_512": err_o}
    if base is None:
        base = (W.clone(), D.clone())
    else:
        res["max_abs_vs_first"] = float(max((W - base[0]).abs().max(), (D - base[1]).abs().max()))
    print(json.dumps(res), flush=True)
