"""One lfb_qr_tsqr_dev_f64 call (for an ncu launch list of the tall-skinny route)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linfa_linalg_b200 as L

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
e = L.engine()
A = torch.rand((cols, rows), dtype=torch.float64, device="cuda") * 2 - 1
d = torch.empty(cols, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
e.set_stream(torch.cuda.current_stream().cuda_stream)
e._check(e.call("lfb_qr_tsqr_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(d.data_ptr())))
torch.cuda.synchronize()
print("launches", e.launch_count)
