import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "fasttanhsinhquadrature.jl_b200"))
import fts_b200 as fts
fts.init(0)
warnings.simplefilter("ignore")
def best(fn, reps=3):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r
c3 = fts.adaptive_cache_3D(np.float64, max_levels=6)
c2 = fts.adaptive_cache_2D(np.float64, max_levels=8)
for L in (4, 5, 6):
    t, r = best(lambda: fts.adaptive_integrate_3D(np.float64, fts.builtin("log_3d"), [-1.0] * 3, [1.0] * 3, rtol=0.0, atol=0.0, max_levels=L, cache=c3, order="reference", full_output=True))
    print("ORD", os.environ.get("FTS_B200_ORDERED", "1"), "3D log_3d single L", L, "level", int(r.levels[0]), "ms", round(t, 3), "value", repr(float(r.value[0])))
t, r = best(lambda: fts.adaptive_integrate_2D(np.float64, fts.builtin("invsqrtabs_2d"), [-1.0, -1.0], [2.0, 2.0], rtol=1e-12, max_levels=8, cache=c2, order="reference", full_output=True))
print("ORD", os.environ.get("FTS_B200_ORDERED", "1"), "2D invsqrtabs single L8 level", int(r.levels[0]), "ms", round(t, 3), "value", repr(float(r.value[0])))
rng = np.random.default_rng(0)
n = 4096
low = np.stack([-3 + 2 * rng.random(n), -3 + 2 * rng.random(n)], axis=1); up = np.stack([1 + 2 * rng.random(n), 1 + 2 * rng.random(n)], axis=1)
cc = fts.adaptive_cache_2D_cmpl(np.float64, max_levels=8)
t, r = best(lambda: fts.quad_cmpl(fts.builtin("c2_invsqrt_bmx_ymc"), low, up, rtol=1e-10, max_levels=8, cache=cc, order="reference", full_output=True))
print("ORD", os.environ.get("FTS_B200_ORDERED", "1"), "config C reference order 4096 boxes ms", round(t, 3), "sum", repr(float(r.value.sum())))
