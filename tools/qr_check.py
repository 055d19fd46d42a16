"""Device-resident QR correctness check at any size (no oracle): ||A - QR|| via R^T R = A^T A and diag/β signs.
usage: python tools/qr_check.py n m [opt=val ...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
eng = L.Engine(0)
for k, v in (a.split("=") for a in sys.argv[3:]):
    eng.set_option(k, int(v))
eng.set_stream(torch.cuda.current_stream().cuda_stream)
A = torch.rand((n, m), dtype=torch.float64, device="cuda") * 2 - 1      # column-major m x n
W = A.clone()
d = torch.zeros(n, dtype=torch.float64, device="cuda")
st = eng.lib.lfb_qr_dev_f64(eng.h, C.c_void_p(W.data_ptr()), m, n, m, C.c_void_p(d.data_ptr()))
assert st == 0
torch.cuda.synchronize()
# R (n x n upper, column-major) from the compact factor: torch view W[c, r] = element (r, c)
Rt = W[:, :n].clone()                      # Rt[c, r] = QR[r, c], r < n
R = torch.triu(Rt.t(), 1) + torch.diag(d.abs())
AtA = A @ A.t()                            # (n x n) = A_math^T A_math  since A tensor is A_math^T
err = (R.t() @ R - AtA).norm() / AtA.norm()
print(f"n={n} m={m}: ||R^T R - A^T A|| / ||A^T A|| = {err:.3e}")
assert err < 1e-12
