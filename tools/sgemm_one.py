"""One TN f32 GEMM on the tcgen05 3xTF32 kernel (for ncu): python tools/sgemm_one.py M N K [reps] [sgemm_tc]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

M, N, K = [int(x) for x in sys.argv[1:4]]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = L.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
if len(sys.argv) > 5:
    eng.set_option("sgemm_tc", int(sys.argv[5]))
A = torch.rand((M, K), dtype=torch.float32, device="cuda") - 0.5
B = torch.rand((N, K), dtype=torch.float32, device="cuda") - 0.5
Cm = torch.zeros((N, M), dtype=torch.float32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    st = eng.lib.lfb_gemm_dev_f32(eng.h, 1, 0, M, N, K, 1.0, C.c_void_p(A.data_ptr()), K, C.c_void_p(B.data_ptr()), K, 0.0,
                                  C.c_void_p(Cm.data_ptr()), M)
    assert st == 0
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"TN f32 {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.2f} TFLOP/s")
