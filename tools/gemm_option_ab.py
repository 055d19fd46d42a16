"""A 0/1 option of the TMA DGEMM (default gemm_tma2: 128 x 64 tiles, two CTAs per SM) against the 128 x 128 one-CTA-per-SM kernel
on the shapes the factorisations produce: bitwise-equal results (same k order, same epilogue arithmetic) and CUDA-event times.
usage: python tools/gemm_persist.py [option]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

OPT = sys.argv[1] if len(sys.argv) > 1 else "gemm_tma2"      # "name" or "name:on_value"
ON = 1
if ":" in OPT:
    OPT, on = OPT.split(":")
    ON = int(on)
eng = L.Engine(0)
for kv in sys.argv[2:]:                                        # extra options held fixed for both arms
    k, v = kv.split("=")
    eng.set_option(k, int(v))
lib = eng.lib
eng.set_stream(torch.cuda.current_stream().cuda_stream)
shapes = [  # (ta, tb, M, N, K, alpha, beta, tag)
    (0, 0, 16384, 16256, 128, -1.0, 1.0, "QR C -= V W (K=128)"), (0, 0, 8192, 8064, 128, -1.0, 1.0, "QR C -= V W half"),
    (0, 0, 16384, 16128, 256, -1.0, 1.0, "QR nb=256"), (1, 0, 8192, 8192, 512, -1.0, 1.0, "SYRK TN K=512"),
    (1, 0, 15872, 15872, 512, -1.0, 1.0, "SYRK TN K=512 full"), (0, 1, 8192, 8192, 512, -1.0, 1.0, "NT K=512"),
    (0, 0, 8192, 8192, 8192, 1.0, 0.0, "square NN"), (1, 0, 8192, 8192, 8192, 1.0, 0.0, "square TN"),
    (1, 0, 128, 16256, 16384, 1.0, 0.0, "QR W = V^T C (split-K)"), (1, 0, 128, 8064, 8192, 1.0, 0.0, "QR W = V^T C half (split-K)"),
    (1, 0, 256, 16128, 16384, 1.0, 0.0, "QR W nb=256 (split-K)"),
    (1, 1, 4096, 4096, 1024, 0.5, 0.0, "TT"), (0, 0, 5003, 3001, 130, -1.0, 1.0, "ragged NN (odd M: scalar-C path, not persistent)"),
    (0, 0, 5002, 3001, 130, -1.0, 1.0, "ragged NN"), (1, 0, 2050, 2306, 770, 2.0, 1.0, "ragged TN"), (0, 1, 1990, 2110, 64, 1.0, 0.5, "ragged NT"),
]
for ta, tb, M, N, K, alpha, beta, tag in shapes:
    A = torch.rand((M, K) if not ta else (K, M), dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand((K, N) if not tb else (N, K), dtype=torch.float64, device="cuda") - 0.5
    C0 = torch.rand((N, M), dtype=torch.float64, device="cuda") - 0.5
    Acm, Bcm = A.t().contiguous(), B.t().contiguous()
    out, tf = {}, {}
    for pers in (0, 1):
        eng.set_option(OPT, ON if pers else 0)
        Cm = C0.clone()

        def run():
            st = lib.lfb_gemm_dev_f64(eng.h, ta, tb, M, N, K, alpha, C.c_void_p(Acm.data_ptr()), A.shape[0], C.c_void_p(Bcm.data_ptr()),
                                      B.shape[0], beta, C.c_void_p(Cm.data_ptr()), M)
            assert st == 0
        run()
        out[pers] = Cm.clone()
        run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        tf[pers] = 2.0 * M * N * K / (e0.elapsed_time(e1) / reps * 1e-3) / 1e12
    ref = alpha * ((A.t() if ta else A) @ (B.t() if tb else B)) + beta * C0.t()
    print(json.dumps({"shape": tag, "ta": ta, "tb": tb, "M": M, "N": N, "K": K, "off_tflops": round(tf[0], 2), "on_tflops": round(tf[1], 2),
                      "bitwise_equal": bool(torch.equal(out[0], out[1])), "max_err_vs_cublas": float((out[1].t() - ref).abs().max())}), flush=True)
    del A, B, C0, Acm, Bcm, out, ref, Cm
    torch.cuda.empty_cache()
