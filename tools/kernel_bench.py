"""Times single internal kernels back to back (CUDA events inside the library).
usage: python tools/kernel_bench.py name n [reps] [opt=value ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 and "=" not in sys.argv[3] else 200
eng = L.Engine(0)
for a in sys.argv[3:]:
    if "=" in a:
        k, v = a.split("=")
        eng.set_option(k, int(v))
us = C.c_double(0)
st = eng.lib.lfb_microbench_kernel(eng.h, name.encode(), n, reps, C.byref(us))
assert st == 0, eng.lib.lfb_last_error(eng.h)
print(f"{name} n={n} {' '.join(a for a in sys.argv[3:] if '=' in a)}: {us.value:.2f} us per launch"
      + (f", {4.0 * n * n / us.value / 1e3:.0f} GB/s of lower-triangle traffic" if name == "trd_symv" else "")
      + (f", {8.0 * 4 * n * n / us.value / 1e3:.0f} GB/s" if name.startswith("bd_gemv") else ""))
