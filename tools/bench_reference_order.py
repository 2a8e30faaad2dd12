#!/usr/bin/env python
"""Wall time of ONE (or a few) 2D/3D integral(s) in the reference summation order — the path the
ordered cooperative kernels serve (k_adaptive{2,3}d_ord, k_fixed{1,2,3}d_ord).  Run it twice to
compare with one thread per integral:

    python tools/bench_reference_order.py
    FTS_B200_ORDERED=0 python tools/bench_reference_order.py
"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fasttanhsinhquadrature.jl_b200"))
import fts_b200 as fts  # noqa: E402


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r


def main():
    fts.init(0)
    warnings.simplefilter("ignore")
    mode = "thread per integral" if os.environ.get("FTS_B200_ORDERED") == "0" else "ordered CTA per integral"
    c3 = fts.adaptive_cache_3D(np.float64, max_levels=6)
    c2 = fts.adaptive_cache_2D(np.float64, max_levels=8)
    for L in (4, 5, 6):
        t, r = best(lambda: fts.adaptive_integrate_3D(np.float64, fts.builtin("log_3d"), [-1.0] * 3, [1.0] * 3, rtol=0.0, atol=0.0, max_levels=L,
                                                      cache=c3, order="reference", full_output=True))
        print(json.dumps(dict(case="adaptive_integrate_3D log_3d, one integral", mode=mode, level=int(r.levels[0]), ms=t, value=float(r.value[0]))))
    t, r = best(lambda: fts.adaptive_integrate_2D(np.float64, fts.builtin("invsqrtabs_2d"), [-1.0, -1.0], [2.0, 2.0], rtol=1e-12, max_levels=8,
                                                  cache=c2, order="reference", full_output=True))
    print(json.dumps(dict(case="adaptive_integrate_2D invsqrtabs_2d, one integral", mode=mode, level=int(r.levels[0]), ms=t, value=float(r.value[0]))))
    rng = np.random.default_rng(0)
    n = 4096
    low = np.stack([-3 + 2 * rng.random(n), -3 + 2 * rng.random(n)], axis=1)
    up = np.stack([1 + 2 * rng.random(n), 1 + 2 * rng.random(n)], axis=1)
    cc = fts.adaptive_cache_2D_cmpl(np.float64, max_levels=8)
    t, r = best(lambda: fts.quad_cmpl(fts.builtin("c2_invsqrt_bmx_ymc"), low, up, rtol=1e-10, max_levels=8, cache=cc, order="reference", full_output=True))
    print(json.dumps(dict(case="config C (4096 boxes) in reference order", mode=mode, ms=t, checksum=float(r.value.sum()))))
    for N in (64, 128, 256):
        x, w, h = fts.tanhsinh(np.float64, N)
        t, r = best(lambda: fts.integrate3D(fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, x, w, h, order="reference"))
        print(json.dumps(dict(case="integrate3D exp_3d, one integral", mode=mode, N=N, ms=t, value=float(r))))
    x, w, h = fts.tanhsinh(np.float64, 2048)
    t, r = best(lambda: fts.integrate2D(fts.builtin("exp_2d"), [-1.0] * 2, [1.0] * 2, x, w, h, order="reference"))
    print(json.dumps(dict(case="integrate2D exp_2d, one integral", mode=mode, N=2048, ms=t, value=float(r))))


if __name__ == "__main__":
    main()
