"""One engine FP64 GEMM launch (for ncu): python tools/gemm_one.py ta tb M N K [reps] [gemm_tma]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

ta, tb, M, N, K = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
eng = L.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
if len(sys.argv) > 7:
    eng.set_option("gemm_tma", int(sys.argv[7]))
A = torch.rand((K, M) if not ta else (M, K), dtype=torch.float64, device="cuda") - 0.5   # column-major storage
B = torch.rand((N, K) if not tb else (K, N), dtype=torch.float64, device="cuda") - 0.5
Cm = torch.zeros((N, M), dtype=torch.float64, device="cuda")
lda = M if not ta else K
ldb = K if not tb else N
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    st = eng.lib.lfb_gemm_dev_f64(eng.h, ta, tb, M, N, K, 1.0, C.c_void_p(A.data_ptr()), lda, C.c_void_p(B.data_ptr()), ldb, 0.0,
                                  C.c_void_p(Cm.data_ptr()), M)
    assert st == 0
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{ta}{tb} {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.2f} TFLOP/s")
