"""Device time of the tall-skinny leaf on ONE GPU for the shard sizes of an 8-way row split (what bounds the scaling).
usage: python tools/tsqr_shard_time.py [cols]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = L.Engine(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
eng.set_stream(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
for rows in (4194304, 2097152, 1048576, 524288, 2048):
    A0 = torch.rand((cols, rows), dtype=torch.float64, device="cuda") * 2 - 1
    A = torch.empty_like(A0)
    R = torch.zeros((cols, cols), dtype=torch.float64, device="cuda")
    d = torch.zeros(cols, dtype=torch.float64, device="cuda")
    out = {"rows": rows, "cols": cols}
    for name, fn in (("tsqr_r", lambda: eng.lib.lfb_tsqr_local_r_dev_f64(eng.h, p(A), rows, cols, rows, p(R), cols)),
                     ("qr_tsqr", lambda: eng.lib.lfb_qr_tsqr_dev_f64(eng.h, p(A), rows, cols, rows, p(d)))):
        best = 1e30
        for it in range(5):
            A.copy_(A0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = eng.launch_count
            e0.record(); st = fn(); e1.record()
            torch.cuda.synchronize()
            assert st == 0
            if it > 0:
                best = min(best, e0.elapsed_time(e1))
        out[name + "_ms"] = round(best, 3)
        out[name + "_launches"] = eng.launch_count - l0
    print(json.dumps(out), flush=True)
    del A0, A
    torch.cuda.empty_cache()
