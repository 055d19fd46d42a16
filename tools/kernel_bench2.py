"""Per-launch device times of small internal kernels (lfb_microbench_kernel).  usage: python tools/kernel_bench2.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L  # noqa: E402

eng = L.Engine(0)
for name, n in (("potf2", 0), ("potf2", 1), ("potf2", 2), ("potf2", 3)):
    us = C.c_double(0)
    st = eng.lib.lfb_microbench_kernel(eng.h, name.encode(), n, 200, C.byref(us))
    print(f"{name} variant {n}: status {st}, {us.value:.2f} us per launch", flush=True)
