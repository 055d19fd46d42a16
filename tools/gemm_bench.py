"""Times the engine's FP64 GEMM on the shapes the factorisations produce, next to cuBLAS (torch.matmul,
yard-stick only: the product never links cuBLAS).  Run on the GPU box: python tools/gemm_bench.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402


def main():
    eng = L.Engine(0)
    lib = eng.lib
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    shapes = [  # (ta, tb, M, N, K, tag)
        (0, 0, 8192, 8192, 8192, "square NN"), (1, 0, 8192, 8192, 8192, "square TN"), (0, 1, 8192, 8192, 8192, "square NT"),
        (0, 0, 4096, 4096, 4096, "square NN 4096"),
        (1, 0, 128, 16384, 16384, "QR W=V^T C (skinny out, long K)"), (0, 0, 16384, 16384, 128, "QR C-=V W (K=128)"),
        (1, 0, 128, 8192, 8192, "QR W=V^T C half"), (0, 0, 8192, 8192, 128, "QR C-=V W half"),
        (0, 1, 8192, 8192, 512, "SYRK-like NT K=512"), (0, 1, 8192, 8192, 2048, "NT K=2048"), (0, 1, 8192, 64, 64, "tall NT K=64"),
        (1, 0, 32, 32, 16384, "Gram 32x32"), (1, 0, 128, 128, 16384, "Gram 128x128"),
    ]
    out = []
    for ta, tb, M, N, K, tag in shapes:
        A = torch.rand((M, K) if not ta else (K, M), dtype=torch.float64, device="cuda") - 0.5
        B = torch.rand((K, N) if not tb else (N, K), dtype=torch.float64, device="cuda") - 0.5
        Cm = torch.zeros((N, M), dtype=torch.float64, device="cuda")
        Acm, Bcm = A.t().contiguous(), B.t().contiguous()
        res = {}
        for tma in (0, 1):
            eng.set_option("gemm_tma", tma)
            def run():
                st = lib.lfb_gemm_dev_f64(eng.h, ta, tb, M, N, K, 1.0, C.c_void_p(Acm.data_ptr()), A.shape[0], C.c_void_p(Bcm.data_ptr()),
                                          B.shape[0], 0.0, C.c_void_p(Cm.data_ptr()), M)
                assert st == 0
            for _ in range(2):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res["tma" if tma else "ldg"] = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        Aop = A.t() if ta else A
        Bop = B.t() if tb else B
        ref = Aop @ Bop
        err = float((Cm.t() - ref).abs().max())
        for _ in range(2):
            Aop @ Bop
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            Aop @ Bop
        e1.record()
        torch.cuda.synchronize()
        cublas = 2.0 * M * N * K / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12
        row = {"shape": tag, "M": M, "N": N, "K": K, "ta": ta, "tb": tb, "ldg_tflops": round(res["ldg"], 2), "tma_tflops": round(res["tma"], 2),
               "cublas_tflops": round(cublas, 2), "max_err": err}
        print(json.dumps(row), flush=True)
        out.append(row)
        del A, B, Cm, Acm, Bcm, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
