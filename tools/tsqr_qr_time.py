"""Times qr_into of a tall-skinny f64 matrix three ways on one GPU: blocked compact-WY (lfb_qr_dev_f64), R-only TSQR
(lfb_tsqr_local_r_dev_f64) and TSQR + Householder reconstruction (lfb_qr_tsqr_dev_f64, same output as the first)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linfa_linalg_b200 as L

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
e = L.engine()
dev = torch.device("cuda:0")
A0 = torch.rand((cols, rows), dtype=torch.float64, device=dev) * 2 - 1
A = torch.empty_like(A0)
d = torch.empty(cols, dtype=torch.float64, device=dev)
R = torch.empty((cols, cols), dtype=torch.float64, device=dev)
s = torch.cuda.current_stream()
e.set_stream(s.cuda_stream)


def run(name, *args):
    def f():
        A.copy_(A0)
        e._check(e.call(name, C.c_void_p(A.data_ptr()), rows, cols, rows, *args))
    f(); f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3):
        f()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3


out = {"rows": rows, "cols": cols}
out["tsqr_r_only_ms"] = run("lfb_tsqr_local_r_dev_f64", C.c_void_p(R.data_ptr()), cols)
out["qr_tsqr_hr_ms"] = run("lfb_qr_tsqr_dev_f64", C.c_void_p(d.data_ptr()))
F2 = A.clone(); d2 = d.clone()
out["qr_blocked_ms"] = run("lfb_qr_dev_f64", C.c_void_p(d.data_ptr()))
out["factor_max_diff_vs_blocked"] = float((A - F2).abs().max())
out["diag_max_diff_vs_blocked"] = float((d - d2).abs().max())
# stage breakdown of the TSQR + reconstruction route through its building-block entry points
U = torch.empty((cols, cols), dtype=torch.float64, device=dev)


def stage(fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        e0.record(s)
        fn()
        e1.record(s)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def st_q():
    e._check(e.call("lfb_tsqr_explicit_q_dev_f64", C.c_void_p(A.data_ptr()), rows, cols, rows, C.c_void_p(R.data_ptr()), cols))


def st_top():
    e._check(e.call("lfb_hh_reconstruct_top_dev_f64", C.c_void_p(A.data_ptr()), cols, rows, C.c_void_p(R.data_ptr()), cols,
                    C.c_void_p(U.data_ptr()), cols, C.c_void_p(d.data_ptr())))


def st_rows():
    e._check(e.call("lfb_hh_reconstruct_rows_dev_f64", C.c_void_p(A.data_ptr() + cols * 8), rows - cols, cols, rows,
                    C.c_void_p(U.data_ptr()), cols))


A.copy_(A0)
out["stage_explicit_q_ms"] = stage(st_q, 1)
out["stage_reconstruct_top_ms"] = stage(st_top, 1)
out["stage_reconstruct_rows_ms"] = stage(st_rows, 1)
fl = 2.0 * rows * cols * cols - 2.0 / 3.0 * cols ** 3
out["qr_tsqr_hr_gflops_equiv"] = fl / (out["qr_tsqr_hr_ms"] * 1e-3) / 1e9
out["qr_blocked_gflops"] = fl / (out["qr_blocked_ms"] * 1e-3) / 1e9
print(json.dumps(out))
