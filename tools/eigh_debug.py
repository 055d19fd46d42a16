import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_linalg_b200 as L
n = int(sys.argv[1]); dt = np.float32 if sys.argv[2] == "f32" else np.float64
g = np.random.default_rng(n).uniform(-100, 100, (n, n))
a0 = ((g + g.T) / 2).astype(dt)
eng = L.Engine(0)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
for it in range(int(sys.argv[3])):
    vals, vecs = L.eigh(a0, eng)
    q = vecs.astype(np.float64)
    res = np.linalg.norm(a0.astype(np.float64) @ q - q * vals.astype(np.float64)[None, :])
    d = L.sym_tridiagonal(a0.copy(), eng)
    qq = d.generate_q(); tt = d.into_tridiag_matrix()
    print(it, "resid", res, "orth", np.linalg.norm(q.T @ q - np.eye(n)), "tridiag resid", np.linalg.norm(qq @ tt @ qq.T - a0))
