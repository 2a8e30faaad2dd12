import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"))
import numpy as np, mpmath
import fts_b200 as fts
mpmath.mp.dps = 40
fts.init(0)
x, w, h = fts.tanhsinh(fts.Double64, 80)
print("tanhsinh dd h", h, x[:2], w[:2])
v = fts.integrate1D(fts.builtin("exp"), fts.to_dd([0.0]), fts.to_dd([1.0]), x, w, h)
print("fixed exp dd", v, fts.dd_to_mpf(v)[0] - (mpmath.e - 1))
c = fts.adaptive_cache_1D(fts.Double64, max_levels=10)
r = fts.adaptive_integrate_1D(fts.Double64, fts.builtin("exp"), fts.to_dd([0.0]), fts.to_dd([1.0]), rtol="1e-25", max_levels=10, cache=c, full_output=True, warn=False)
print("adaptive exp dd", r, fts.dd_to_mpf(r.value)[0] - (mpmath.e - 1))
r = fts.adaptive_integrate_1D(fts.Double64, fts.builtin("poly7"), fts.to_dd([0.0]), fts.to_dd([1.0]), rtol="1e-25", max_levels=10, cache=c, full_output=True, warn=False, params=fts.to_dd([1,1,0,0,0,0,0,0]).reshape(1,8))
print("adaptive poly dd", r)
cc = fts.adaptive_cache_1D(fts.Double64, max_levels=10, complement=True)
print("cmpl cache tm", cc.tm, "ix", cc.initial_x, "iw", cc.initial_w, "ic", cc.initial_c)
for ml in (0, 1, 2, 10):
    r = fts.adaptive_integrate_1D_cmpl(fts.Double64, fts.builtin("c_invsqrt"), fts.to_dd([-1.0]), fts.to_dd([1.0]), rtol="1e-25", max_levels=ml, cache=cc, full_output=True, warn=False, params=fts.to_dd([1.0]))
    print("cmpl invsqrt ml", ml, r)
r = fts.adaptive_integrate_1D_cmpl(fts.Double64, fts.builtin("c_exp_plus"), fts.to_dd([-1.0]), fts.to_dd([1.0]), rtol="1e-25", max_levels=10, cache=cc, full_output=True, warn=False)
print("cmpl exp_plus", r)
r = fts.adaptive_integrate_1D_cmpl(fts.Double64, fts.builtin("c_log_bmx"), fts.to_dd([-1.0]), fts.to_dd([1.0]), rtol="1e-25", max_levels=10, cache=cc, full_output=True, warn=False)
print("cmpl log", r)
