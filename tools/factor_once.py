"""Runs one warm-up + one timed device-resident factorisation (for ncu launch lists).
usage: python tools/factor_once.py {chol|qr|tsqr|tridiag|bidiag|batched} n [m]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

kind = sys.argv[1]
n = int(sys.argv[2])
m = int(sys.argv[3]) if len(sys.argv) > 3 else n
eng = L.Engine(0)
lib = eng.lib
eng.set_stream(torch.cuda.current_stream().cuda_stream)
warm = 1
for k, v in (a.split("=") for a in sys.argv[4:]):
    if k == "warm":
        warm = int(v)
    else:
        eng.set_option(k, int(v))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
p = lambda t: C.c_void_p(t.data_ptr())
if kind == "chol":
    S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    S = (S + S.t()) / 2
    S.diagonal().add_(float(n))
    W = torch.empty_like(S)
    info = torch.zeros(1, dtype=torch.int64, device=dev)
    step = lambda: (W.copy_(S), lib.lfb_cholesky_dev_f64(eng.h, p(W), n, n, 0, p(info)))
    flops = n ** 3 / 3
elif kind == "qrf":
    A = torch.rand((n, m), dtype=torch.float32, device=dev, generator=g) * 2 - 1   # column-major m x n
    W = torch.empty_like(A)
    d = torch.zeros(n, dtype=torch.float32, device=dev)
    step = lambda: (W.copy_(A), lib.lfb_qr_dev_f32(eng.h, p(W), m, n, m, p(d)))
    flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
elif kind in ("qr", "tsqr"):
    A = torch.rand((n, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1   # column-major m x n
    W = torch.empty_like(A)
    d = torch.zeros(n, dtype=torch.float64, device=dev)
    R = torch.zeros((n, n), dtype=torch.float64, device=dev)
    if kind == "qr":
        step = lambda: (W.copy_(A), lib.lfb_qr_dev_f64(eng.h, p(W), m, n, m, p(d)))
    else:
        step = lambda: (W.copy_(A), lib.lfb_tsqr_local_r_dev_f64(eng.h, p(W), m, n, m, p(R), n))
    flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
elif kind == "tridiag":
    S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    S = (S + S.t()) / 2
    W = torch.empty_like(S)
    off = torch.zeros(n, dtype=torch.float64, device=dev)
    step = lambda: (W.copy_(S), lib.lfb_sym_tridiagonal_dev_f64(eng.h, p(W), n, n, p(off)))
    flops = 4.0 / 3.0 * n ** 3
elif kind == "eigvalsh":
    import numpy as np
    S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    S = (S + S.t()) / 2
    W = torch.empty_like(S)
    vals = np.zeros(n)
    vp = C.c_void_p(vals.ctypes.data)
    step = lambda: (W.copy_(S), lib.lfb_eigh_dev_f64(eng.h, p(W), n, n, vp, None, n))
    flops = 4.0 / 3.0 * n ** 3
elif kind == "eigh":
    import numpy as np
    S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    S = (S + S.t()) / 2
    W = torch.empty_like(S)
    Q = torch.empty_like(S)
    vals = np.zeros(n)
    vp = C.c_void_p(vals.ctypes.data)
    step = lambda: (W.copy_(S), lib.lfb_eigh_dev_f64(eng.h, p(W), n, n, vp, p(Q), n))
    flops = 4.0 / 3.0 * n ** 3
elif kind == "svd":
    import numpy as np
    A = torch.rand((n, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1   # column-major m x n
    W = torch.empty_like(A)
    U = torch.empty((n, m), dtype=torch.float64, device=dev)                       # column-major m x n
    V = torch.empty((n, n), dtype=torch.float64, device=dev)                       # column-major n x n (V = Vt^T)
    sv = np.zeros(n)
    step = lambda: (W.copy_(A), lib.lfb_svd_dev_f64(eng.h, p(W), m, n, m, C.c_void_p(sv.ctypes.data), p(U), m, p(V), n))
    flops = 4.0 * m * n * n - 4.0 / 3.0 * n ** 3
elif kind == "bidiag":
    A = torch.rand((n, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1   # column-major m x n
    W = torch.empty_like(A)
    d = torch.zeros(n, dtype=torch.float64, device=dev)
    e = torch.zeros(n, dtype=torch.float64, device=dev)
    step = lambda: (W.copy_(A), lib.lfb_bidiagonal_dev_f64(eng.h, p(W), m, n, m, p(d), p(e)))
    flops = 4.0 * m * n * n - 4.0 / 3.0 * n ** 3
elif kind == "cholbatched":
    G = torch.rand((n, 32, 32), dtype=torch.float32, device=dev, generator=g) * 2 - 1
    A = G @ G.transpose(1, 2) + 32 * torch.eye(32, device=dev)[None]
    W = torch.empty_like(A)
    f = torch.zeros(n, dtype=torch.int32, device=dev)
    step = lambda: (W.copy_(A), lib.lfb_cholesky_batched_dev_f32(eng.h, p(W), n, 32, 1, p(f)))
    flops = n * 32 ** 3 / 3.0
elif kind == "batched":
    A = torch.rand((n, 32, 32), dtype=torch.float32, device=dev, generator=g) * 2 - 1
    W = torch.empty_like(A)
    d = torch.zeros((n, 32), dtype=torch.float32, device=dev)
    step = lambda: (W.copy_(A), lib.lfb_qr_batched_dev_f32(eng.h, p(W), n, 32, 32, p(d)))
    flops = n * 43690.7
for _ in range(warm):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = eng.launch_count
e0.record()
step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{kind} n={n} m={m}: {ms:.3f} ms, {flops / ms / 1e9:.2f} TFLOP/s, {eng.launch_count - l0} launches")
