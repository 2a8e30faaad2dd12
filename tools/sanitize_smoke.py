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

# ---- round-2 kernels -----------------------------------------------------------------------------------------------
from fts_b200 import _lib as L  # noqa: E402

sm = fts.device_info()["sm_count"]
# warp-per-integral 2D kernel (TMA-staged node slices, mbarriers) + its queue pass: >= 8 x SMs boxes in the cooperative order
nw = 8 * sm + 40
rng = np.random.default_rng(1)
lw = np.stack([-1 - rng.random(nw), -1 - rng.random(nw)], axis=1); uw = np.stack([1 + rng.random(nw), 0.5 + rng.random(nw)], axis=1)
c2c8 = fts.adaptive_cache_2D_cmpl(np.float64, max_levels=6)
print("2d warp cmpl", fts.quad_cmpl(fts.builtin("c2_invsqrt_bmx_ymc"), lw, uw, rtol=1e-8, max_levels=6, cache=c2c8, order="fast")[:2])
print("2d warp f32", fts.quad(fts.builtin("exp_2d"), lw.astype(np.float32), uw.astype(np.float32), rtol=1e-4, max_levels=6,
                                cache=fts.adaptive_cache_2D(np.float32, max_levels=6), order="fast")[:2])
import warnings  # noqa: E402
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    print("2d warp queue", fts.adaptive_integrate_2D_avx(np.float64, fts.builtin("invsqrtabs_2d"), np.full((nw, 2), -1.0), np.full((nw, 2), 2.0), rtol=1e-12,
                                                         max_levels=6, cache=c2)[:2])
# bucket passes (whole batch / incoherent warps only / detection only: FTS_B200_BUCKET = 3 / 2 / 1) on pinned buffers in place
nb = 4 * sm * 256 + 1000
ab = np.exp(np.linspace(np.log(0.5), np.log(3.0e3), nb))[np.random.default_rng(3).permutation(nb)]
lo1, hi1, al1 = fts.pinned_empty(1), fts.pinned_empty(1), fts.pinned_empty(nb)
v1, e1, lv1, fl1 = fts.pinned_empty(nb), fts.pinned_empty(nb), fts.pinned_empty(nb, np.int32), fts.pinned_empty(nb, np.int32)
lo1[...], hi1[...], al1[...] = -1.0, 1.0, ab
c16 = fts.adaptive_cache_1D(np.float64, max_levels=16)
for rep in range(2):  # the second call is the one the feedback of the first arms
    k0 = fts.launch_count()
    fts.adaptive_batch_host(c16, fts.builtin("gauss"), low=lo1.ctypes.data, up=hi1.ctypes.data, bounds_stride=0, params=al1.ctypes.data, params_stride=1,
                            rtol=1e-10, atol=0.0, max_levels=16, batch=nb, value=v1.ctypes.data, err=e1.ctypes.data, levels=lv1.ctypes.data,
                            flags=fl1.ctypes.data, order="reference", options=L.OPT_QUAD_EDGE_RULES)
    print("bucket", os.environ.get("FTS_B200_BUCKET", "1"), "launches", fts.launch_count() - k0, v1[:2], int(lv1.max()))
# Double64: lane groups (1D), thread per integral (2D), quad_split
cdd = fts.adaptive_cache_1D(fts.Double64, max_levels=10, complement=True)
print("dd grp", fts.quad_cmpl(fts.builtin("c_invsqrt"), fts.to_dd(-np.ones(100)), fts.to_dd(np.ones(100)), rtol="1e-26", max_levels=10, cache=cdd,
                              params=fts.to_dd(np.ones(100)))[:1])
cdd2 = fts.adaptive_cache_2D(fts.Double64, max_levels=4)
print("dd 2d", fts.adaptive_integrate_2D(fts.Double64, fts.builtin("exp_2d"), fts.to_dd(lo2[:2]), fts.to_dd(up2[:2]), rtol="1e-20", max_levels=4, cache=cdd2, warn=False)[:1])
print("quad_split", fts.quad_split(fts.builtin("exp_2d"), [0.25, 0.5], [-1.0, -1.0], [1.0, 1.0], rtol=1e-9, max_levels=7, cache=c2))
# fixed-grid complement, generic n-D tensor product, device-built tables, pointwise evaluation
print("integrate1D_cmpl", fts.integrate1D_cmpl(np.float64, fts.builtin("c_invsqrt"), 30, params=[1.0]))
fnd = fts.cuda_integrand("T s = x[0]; for (int d = 1; d < ND; ++d) s = s + x[d]; return exp(s);", "nd", dim=4)
x8, w8, h8 = fts.tanhsinh(np.float64, 8, D=4)
print("nd", fts.integrateND(fnd, [-1.0] * 4, [1.0] * 4, x8, w8, h8))
cdev = fts.adaptive_cache_2D(np.float64, max_levels=5, on_device=True)
print("device tables", fts.adaptive_integrate_2D_avx(np.float64, fts.builtin("exp_2d"), lo2, up2, rtol=1e-8, max_levels=5, cache=cdev)[0])
print("evaluate", fts.evaluate(fts.builtin("gauss"), np.linspace(-1, 1, 5), params=[2.0]))
# reference-order cooperative (cluster / DSMEM) kernels on a deep single integral, fixed-grid ordered kernels at n >= 256
print("ord 2d deep", fts.adaptive_integrate_2D(np.float64, fts.builtin("exp_2d"), [-1.0, -1.0], [1.0, 1.0], rtol=1e-13, max_levels=6, cache=c2, order="reference"))
xl, wl, hl = fts.tanhsinh(np.float64, 300)
print("ord fixed", fts.integrate1D(fts.builtin("exp"), 0.0, 1.0, xl, wl, hl), fts.integrate2D(fts.builtin("exp_2d"), [-1.0, -1.0], [1.0, 1.0], xl, wl, hl))
print("sanitize smoke done")
