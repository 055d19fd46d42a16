"""Option sweeps of one device-resident factorisation in ONE process (min of `reps` timed runs per setting, CUDA events).
usage: python tools/sweep.py {chol|qr|cholf|qrf} n reps "opt=val,opt=val" "opt=val" ...      ("-" = defaults; *f = f32)"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

kind, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
f32 = kind.endswith("f")
kind = kind.rstrip("f")
dt = torch.float32 if f32 else torch.float64
sfx = "_f32" if f32 else "_f64"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
p = lambda t: C.c_void_p(t.data_ptr())
S = torch.rand((n, n), dtype=dt, device=dev, generator=g) * 2 - 1
if kind == "chol":
    S = (S + S.t()) / 2
    S.diagonal().add_(float(n))
W = torch.empty_like(S)
aux = torch.zeros(n, dtype=torch.int64 if kind == "chol" else dt, device=dev)
flops = n ** 3 / 3 if kind == "chol" else 4.0 / 3.0 * n ** 3
for spec in sys.argv[4:]:
    eng = L.Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    prof = False
    if spec != "-":
        for kv in spec.split(","):
            k, v = kv.split("=")
            if k == "prof":                 # serial run under the GEMM profiler (look-ahead off): GEMM share of the factorisation
                prof = int(v) != 0
            else:
                eng.set_option(k, int(v))

    def step():
        W.copy_(S)
        if kind == "chol":
            st = getattr(eng.lib, "lfb_cholesky_dev" + sfx)(eng.h, p(W), n, n, 0, p(aux))
        else:
            st = getattr(eng.lib, "lfb_qr_dev" + sfx)(eng.h, p(W), n, n, n, p(aux))
        assert st == 0
    step(); step()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res = {"kind": kind + ("_f32" if f32 else "_f64"), "n": n, "opts": spec, "ms": round(best, 3), "tflops": round(flops / best / 1e9, 2)}
    if prof:
        gms, gfl, gc = C.c_double(), C.c_double(), C.c_int64()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        W.copy_(S)
        eng.lib.lfb_profile_begin(eng.h)
        e0.record()
        if kind == "chol":
            getattr(eng.lib, "lfb_cholesky_dev" + sfx)(eng.h, p(W), n, n, 0, p(aux))
        else:
            getattr(eng.lib, "lfb_qr_dev" + sfx)(eng.h, p(W), n, n, n, p(aux))
        e1.record()
        torch.cuda.synchronize()
        eng.lib.lfb_profile_end(eng.h, C.byref(gms), C.byref(gfl), C.byref(gc))
        res.update({"prof_total_ms": round(e0.elapsed_time(e1), 3), "prof_gemm_ms": round(gms.value, 3), "prof_gemm_calls": gc.value,
                    "prof_gemm_tflops": round(gfl.value / max(gms.value, 1e-9) / 1e9, 2), "prof_gemm_flop_share": round(gfl.value / flops, 4)})
    if kind == "chol":
        Lf = torch.triu(W[:2048, :2048]).t().double()
        Sd = S[:2048, :2048].double()
        res["resid"] = float((Lf @ Lf.t() - Sd).norm() / Sd.norm())
    else:
        k = min(2048, n)
        Rk = (torch.triu(W[:k, :k].t(), 1) + torch.diag(aux[:k].abs())).double()
        AtA = S[:k, :].double() @ S[:k, :].double().t()
        res["resid"] = float((Rk.t() @ Rk - AtA).norm() / AtA.norm())
    print(json.dumps(res), flush=True)
    eng.set_stream(None)
    eng.close()
