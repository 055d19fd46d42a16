"""TN f32 GEMM (C = A^T B, K-major operands) on the tcgen05 3xTF32 kernel vs single-pass TF32 vs the FFMA kernel vs cuBLAS
(torch, TF32 off).  usage: python tools/sgemm_bench.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
engs = {}
for mode in (1, 2, 0):
    e = L.Engine(0)
    e.set_option("sgemm_tc", mode)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    engs[mode] = e
peak = C.c_double(0)
engs[1].lib.lfb_microbench_fp64(engs[1].h, 2, C.byref(peak))
print(json.dumps({"ffma_peak_tflops": peak.value / 1e3}), flush=True)


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (m, n, k, name) in ((8192, 8192, 8192, "square"), (16384, 16384, 512, "SYRK-like K=512"), (128, 16384, 16384, "QR W=V^T C"),
                        (16384, 16384, 128, "QR C-=V W (K=128)"), (4096, 4096, 4096, "4096^3")):
    A = torch.rand((m, k), dtype=torch.float32, device=dev) - 0.5
    B = torch.rand((n, k), dtype=torch.float32, device=dev) - 0.5
    Cm = torch.zeros((n, m), dtype=torch.float32, device=dev)
    row = {"shape": name, "M": m, "N": n, "K": k}
    fl = 2.0 * m * n * k
    for mode, key in ((1, "tc_3xtf32"), (2, "tc_tf32"), (0, "ffma")):
        if mode == 0 and fl > 3e11:
            continue
        e = engs[mode]
        f = lambda: e.lib.lfb_gemm_dev_f32(e.h, 1, 0, m, n, k, 1.0, C.c_void_p(A.data_ptr()), k, C.c_void_p(B.data_ptr()), k, 0.0,
                                           C.c_void_p(Cm.data_ptr()), m)
        ms = timed(f)
        row[key + "_tflops"] = round(fl / ms / 1e9, 2)
        if mode == 1:
            ref = (A[:256].double() @ B[:256].double().t())
            row["max_err_3xtf32"] = float((Cm.t()[:256, :256].double() - ref).abs().max())
    ms = timed(lambda: torch.matmul(A, B.t()))
    row["cublas_fp32_tflops"] = round(fl / ms / 1e9, 2)
    print(json.dumps(row), flush=True)
