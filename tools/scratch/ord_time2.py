import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "fasttanhsinhquadrature.jl_b200"))
import fts_b200 as fts
fts.init(0)
def best(fn, reps=3):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r
for N in (64, 128, 256):
    x, w, h = fts.tanhsinh(np.float64, N)
    t, r = best(lambda: fts.integrate3D(fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, x, w, h, order="reference"))
    print("ORD", os.environ.get("FTS_B200_ORDERED", "1"), "integrate3D N", N, "ms", round(t, 3), repr(float(r)))
x, w, h = fts.tanhsinh(np.float64, 2048)
t, r = best(lambda: fts.integrate2D(fts.builtin("exp_2d"), [-1.0] * 2, [1.0] * 2, x, w, h, order="reference"))
print("ORD", os.environ.get("FTS_B200_ORDERED", "1"), "integrate2D N 2048 ms", round(t, 3), repr(float(r)))
