#!/usr/bin/env python
"""Small run that touches every kernel style once (for compute-sanitizer memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"))
import numpy as np  # noqa: E402

import fts_b200 as fts  # noqa: E402
from fts_b200 import distributed as fd  # noqa: E402

fts.init(0)
a = 0.5 + 49.5 * np.arange(600) / 599
c1 = fts.adaptive_cache_1D(np.float64, max_levels=12)
print("tpi 1d", fts.quad(fts.builtin("gauss"), -1.0, 1.0, rtol=1e-10, max_levels=12, cache=c1, params=a, order="reference")[:2])
print("tail", fts.quad(fts.builtin("runge"), -1.0, 1.0, rtol=1e-10, max_levels=12, cache=c1, params=np.array([25.0, 1e6, 1e7]), order="reference"))
print("cta 1d", fts.adaptive_integrate_1D_avx(np.float64, fts.builtin("gauss"), -1.0, 1.0, rtol=1e-10, max_levels=12, cache=c1, params=a[:8])[:2])
cc = fts.adaptive_cache_1D(np.float64, max_levels=10, complement=True)
print("cmpl", fts.quad_cmpl(fts.builtin("c_invsqrt"), -1.0, 1.0, rtol=1e-10, max_levels=10, cache=cc, params=[1.0]),
      fts.adaptive_integrate_1D_cmpl_avx(np.float64, fts.builtin("c_invsqrt"), -1.0, 1.0, rtol=1e-10, max_levels=10, cache=cc, params=[1.0]))
x, w, h = fts.tanhsinh(np.float64, 40)
for order in ("reference", "fast"):
    print("fixed", order, fts.integrate1D(fts.builtin("exp"), 0.0, 1.0, x, w, h, order=order), fts.integrate2D(fts.builtin("exp_2d"), [-1.0, -1.0], [1.0, 1.0], x, w, h, order=order),
          fts.integrate3D(fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, x[:10].copy(), w[:10].copy(), h, order=order))
c2 = fts.adaptive_cache_2D(np.float64, max_levels=7)
c3 = fts.adaptive_cache_3D(np.float64, max_levels=4)
lo2 = np.tile([-1.0, -1.0], (5, 1)); up2 = np.tile([1.0, 1.0], (5, 1))
print("2d tpi/cta/grid", fts.adaptive_integrate_2D(np.float64, fts.builtin("exp_2d"), lo2, up2, rtol=1e-8, max_levels=5, cache=c2, order="reference")[0],
      fts.adaptive_integrate_2D_avx(np.float64, fts.builtin("exp_2d"), lo2, up2, rtol=1e-8, max_levels=5, cache=c2)[0],
      fts.adaptive_integrate_2D_avx(np.float64, fts.builtin("exp_2d"), [-1.0, -1.0], [1.0, 1.0], rtol=1e-12, max_levels=7, cache=c2))
c2c = fts.adaptive_cache_2D_cmpl(np.float64, max_levels=5)
print("2d cmpl", fts.quad_cmpl(fts.builtin("c2_invsqrt_bmx_ymc"), lo2, up2, rtol=1e-8, max_levels=5, cache=c2c, order="fast")[0])
lo3 = np.tile([-1.0] * 3, (3, 1)); up3 = np.tile([1.0] * 3, (3, 1))
print("3d tpi/cta/grid", fts.adaptive_integrate_3D(np.float64, fts.builtin("exp_3d"), lo3, up3, rtol=1e-6, max_levels=3, cache=c3, order="reference")[0],
      fts.adaptive_integrate_3D_avx(np.float64, fts.builtin("exp_3d"), lo3, up3, rtol=1e-6, max_levels=3, cache=c3)[0],
      fts.adaptive_integrate_3D_avx(np.float64, fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, rtol=1e-9, max_levels=4, cache=c3))
print("fused", fd.adaptive_integrate_sharded(np.float64, fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, rtol=1e-9, max_levels=4, cache=c3, shard_from_level=0).value)
cd = fts.adaptive_cache_1D(fts.Double64, max_levels=9, complement=True)
print("dd", fts.quad_cmpl(fts.builtin("c_log_bmx"), fts.to_dd([-1.0, -2.0]), fts.to_dd([1.0, 3.0]), rtol="1e-26", max_levels=9, cache=cd))
u = fts.cuda_integrand("return exp(-p[0]*x*x);", "1d", nparams=1)
print("nvrtc", fts.quad(u, -1.0, 1.0, rtol=1e-10, max_levels=12, cache=c1, params=a[:64], order="reference")[:2])
c32 = fts.adaptive_cache_1D(np.float32, max_levels=8)
print("f32", fts.quad(fts.builtin("exp"), np.float32(0), np.float32(1), rtol=1e-5, max_levels=8, cache=c32))
print("sanitize smoke done")
