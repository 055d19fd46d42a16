"""Debug dump of the TSQR + Householder reconstruction stages on a small degenerate input."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import linfa_linalg_b200 as L
import oracle as O

np.set_printoptions(linewidth=200, precision=4, suppress=True)
kind = sys.argv[1] if len(sys.argv) > 1 else "diag"
rows, cols = 3000, 6
a0 = np.random.default_rng(3).uniform(-1, 1, (rows, cols))
if kind == "diag":
    a0[:] = 0
    for k in range(6):
        a0[k, k] = (k + 1) * (-1) ** k
elif kind == "zero_col":
    a0[:, 2] = 0
e = L.Engine(0)
e.set_option("tsqr_chunk", 256)
s = torch.cuda.current_stream()
e.set_stream(s.cuda_stream)
A = torch.from_numpy(np.ascontiguousarray(a0.T)).cuda()
R = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
e._check(e.call("lfb_tsqr_explicit_q_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(R.data_ptr()), cols))
torch.cuda.synchronize()
q = A.t().cpu().numpy().copy(); r = R.t().cpu().numpy().copy()
print("Q top\n", q[:cols], "\nmax|Q rest|", np.abs(q[cols:]).max(), "\nR\n", r)
print("orth", np.linalg.norm(q.T @ q - np.eye(cols)), "QR-A", np.linalg.norm(q @ r - a0))
U = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
d = torch.zeros(cols, dtype=torch.float64, device="cuda")
e._check(e.call("lfb_hh_reconstruct_top_dev_f64", C.c_void_p(A.data_ptr()), cols, rows, C.c_void_p(R.data_ptr()), cols,
                C.c_void_p(U.data_ptr()), cols, C.c_void_p(d.data_ptr())))
torch.cuda.synchronize()
print("top after reconstruct\n", A.t()[:cols].cpu().numpy(), "\nU'\n", U.t().cpu().numpy(), "\ndiag", d.cpu().numpy())
e._check(e.call("lfb_hh_reconstruct_rows_dev_f64", C.c_void_p(A.data_ptr() + cols * 8), rows - cols, cols, rows, C.c_void_p(U.data_ptr()), cols))
torch.cuda.synchronize()
f = A.t().cpu().numpy()
ref = a0.copy(); dref = O.qr(ref)
print("ref top\n", ref[:cols], "\nref diag", dref)
print("max diff lower", np.abs(np.tril(f) - np.tril(ref)).max(), "upper", np.abs(np.triu(f, 1) - np.triu(ref, 1)).max())
# host route
a = a0.copy()
dec = L.qr_tsqr_into(a, eng=e)
print("host route top\n", a[:cols], "\ndiag", dec.diag)
qq, rr = dec.into_decomp()
print("host consumers: QR-A", np.linalg.norm(qq @ rr - a0), "orth", np.linalg.norm(qq.T @ qq - np.eye(cols)))
a = a0.copy()
dec = L.qr_into(a, eng=e)
qq, rr = dec.into_decomp()
print("blocked route diag", dec.diag, np.signbit(dec.diag), "QR-A", np.linalg.norm(qq @ rr - a0), "orth", np.linalg.norm(qq.T @ qq - np.eye(cols)))
