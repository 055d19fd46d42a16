"""Per-phase cycle counts of one cluster panel launch (CTA 0): python tools/panel_phases.py m n"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L
m, n = int(sys.argv[1]), int(sys.argv[2])
eng = L.Engine(0)
eng.set_option("lookahead", 0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
out = (C.c_longlong * 4)()
eng.lib.lfb_debug_panel_phases(eng.h, out)          # arm
A = torch.rand((n, m), dtype=torch.float64, device="cuda") * 2 - 1
d = torch.zeros(n, dtype=torch.float64, device="cuda")
eng.lib.lfb_qr_dev_f64(eng.h, C.c_void_p(A.data_ptr()), m, n, m, C.c_void_p(d.data_ptr()))
torch.cuda.synchronize()
eng.lib.lfb_debug_panel_phases(eng.h, out)
w = 32
names = ["inbox sum", "scalars+fac", "row pass", "reduce+push+cluster barrier"]
tot = sum(out)
for nm, v in zip(names, out):
    print(f"{nm:32s} {v / w:9.0f} cycles/column  {100.0 * v / max(tot, 1):5.1f}%")
print(f"total {tot / w:.0f} cycles/column (last sub-panel launch of QR {m}x{n}, rows ~{m - n + 32})")
