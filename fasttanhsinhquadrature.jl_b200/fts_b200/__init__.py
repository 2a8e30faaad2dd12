"""fts_b200 — B200 (sm_100a) implementation of the FastTanhSinhQuadrature.jl hot path.

Python host mirror of the reference's Julia API over the C ABI of libfts_b200.so
(include/fts_b200.h).  See DESIGN.md and INTEGRATION.md at the repository root.
"""
from ._lib import (ArgumentError, DimensionMismatch, FtsError, LIB_PATH, FLAG_CONVERGED, FLAG_FLIPPED, FLAG_MAX_LEVELS,
                   FLAG_ZERO_WIDTH, KIND_1D, KIND_1D_CMPL, KIND_2D, KIND_2D_CMPL, KIND_3D, ORDER_AUTO, ORDER_FAST, ORDER_REFERENCE)
from . import api, distributed
from .api import (FAMILIES, AdaptiveCache, AdaptiveResult, DeviceTables, ConvergenceWarning, Double64, Integrand, adaptive_cache_1D,
                  adaptive_cache_2D, adaptive_cache_2D_cmpl, adaptive_cache_3D, adaptive_integrate_1D, adaptive_integrate_1D_avx,
                  adaptive_integrate_1D_cmpl, adaptive_integrate_1D_cmpl_avx, adaptive_integrate_2D, adaptive_integrate_2D_avx,
                  adaptive_integrate_2D_cmpl, adaptive_integrate_3D, adaptive_integrate_3D_avx, builtin, cuda_integrand, dd_add, dd_to_mpf,
                  device_info, init, integrate1D, integrate1D_avx, integrate1D_cmpl, integrate2D, integrate2D_avx, integrate3D, integrate3D_avx, integrateND,
                  launch_count, measure_fma_peak, evaluate, pinned_empty, adaptive_batch_dev, adaptive_batch_host, evals_for_level, quad, quad_cmpl, quad_split, t_w_max, t_x_max, tanhsinh, tanhsinh_opt, hopt, hmax, tmax, to_dd)

# the 21 names exported by the reference (src/FastTanhSinhQuadrature.jl:44-51)
REFERENCE_EXPORTS = (
    "tanhsinh", "quad", "quad_split", "quad_cmpl", "integrate1D", "integrate2D", "integrate3D", "integrate1D_avx",
    "integrate2D_avx", "integrate3D_avx", "adaptive_integrate_1D", "adaptive_integrate_2D", "adaptive_integrate_3D",
    "adaptive_integrate_1D_cmpl", "adaptive_integrate_1D_avx", "adaptive_integrate_2D_avx", "adaptive_integrate_3D_avx",
    "adaptive_integrate_1D_cmpl_avx", "adaptive_cache_1D", "adaptive_cache_2D", "adaptive_cache_3D")
