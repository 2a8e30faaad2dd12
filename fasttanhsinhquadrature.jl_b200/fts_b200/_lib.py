"""ctypes binding of libfts_b200.so (include/fts_b200.h).

The library is the product: hand-written sm_100a CUDA kernels behind a C ABI.  Importing works on
a CPU-only machine (so host logic and the ABI surface can be tested), but every compute entry
point fails loudly with FtsError(FTS_ERR_NO_DEVICE) when no B200 is present — there is no CPU
fallback, and the repository's CPU test checker is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# FTS_B200_LIB: another build of the same library (experiments with compile-time constants); default = the in-tree build
LIB_PATH = os.environ.get("FTS_B200_LIB") or os.path.join(_ROOT, "lib", "libfts_b200.so")

# status codes / enums (mirror of the header; tests/test_abi.py checks them against it)
OK, ERR_INVALID_ARGUMENT, ERR_DIMENSION_MISMATCH, ERR_UNSUPPORTED, ERR_CUDA, ERR_NVRTC, ERR_NO_DEVICE, ERR_INTERNAL = range(8)
F32, F64, DD64 = 0, 1, 2
ORDER_REFERENCE, ORDER_FAST, ORDER_AUTO = 0, 1, 2
OPT_QUAD_EDGE_RULES = 1
OPT_FORCE_DEPTH = 2
FLAG_CONVERGED, FLAG_MAX_LEVELS, FLAG_ZERO_WIDTH, FLAG_FLIPPED = 1, 2, 4, 8
KIND_1D, KIND_2D, KIND_3D, KIND_1D_CMPL, KIND_2D_CMPL = 1, 2, 3, 4, 5
KIND_ND, ND_MAX = 16, 8  # generic n-D tensor products: kind = KIND_ND + D


class FtsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[fts_b200 status {code}] {message}")
        self.code = code
        self.message = message


class DimensionMismatch(FtsError, ValueError):
    """mirror of Julia's DimensionMismatch (integrate2D.jl:149-150, quad.jl:70)"""


class ArgumentError(FtsError, ValueError):
    """mirror of Julia's ArgumentError (core.jl:41-42, caches.jl:26-27)"""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fts_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

vp, ci, cs_, i64 = C.c_void_p, C.c_int, C.c_size_t, C.c_int64
_SIGS = {
    "fts_abi_version": (ci, []),
    "fts_init": (ci, [ci]),
    "fts_shutdown": (ci, []),
    "fts_device_info": (ci, [C.POINTER(ci)] * 4 + [C.POINTER(cs_)]),
    "fts_last_error": (cs_, [C.c_char_p, cs_]),
    "fts_host_alloc": (ci, [C.POINTER(vp), cs_]),
    "fts_host_free": (ci, [vp]),
    "fts_family_info": (ci, [ci, C.POINTER(ci), C.POINTER(ci)]),
    "fts_window": (ci, [ci, ci, ci, vp]),
    "fts_tanhsinh": (ci, [ci, ci, ci, vp, vp, vp]),
    "fts_hopt": (ci, [ci, ci, C.c_double, vp]),
    "fts_hmax": (ci, [ci, ci, ci, vp]),
    "fts_tanhsinh_opt": (ci, [ci, ci, C.c_double, ci, i64, vp, vp, vp, C.POINTER(i64)]),
    "fts_cache1d_generate": (ci, [ci, ci, ci] + [vp] * 7),
    "fts_cache_tensor_generate": (ci, [ci, ci, ci] + [vp] * 5),
    "fts_cache_tensor_cmpl_generate": (ci, [ci, ci] + [vp] * 7),
    "fts_tables_fixed_upload": (ci, [ci, ci, vp, vp, vp, C.POINTER(vp)]),
    "fts_tables_free": (ci, [vp]),
    "fts_cache1d_upload": (ci, [ci, ci, ci] + [vp] * 7 + [C.POINTER(vp)]),
    "fts_cache_tensor_upload": (ci, [ci, ci, ci] + [vp] * 5 + [C.POINTER(vp)]),
    "fts_cache_tensor_cmpl_upload": (ci, [ci, ci] + [vp] * 7 + [C.POINTER(vp)]),
    "fts_tables_fixed_create": (ci, [ci, ci, ci, C.POINTER(vp)]),
    "fts_cache_create": (ci, [ci, ci, ci, ci, C.POINTER(vp)]),
    "fts_tables_download": (ci, [vp, C.POINTER(ci), vp, vp, vp]),
    "fts_cache_download": (ci, [vp] + [vp] * 7),
    "fts_cache_info": (ci, [vp] + [C.POINTER(ci)] * 4),
    "fts_cache_free": (ci, [vp]),
    "fts_integrand_builtin": (ci, [ci, C.POINTER(vp)]),
    "fts_integrand_nvrtc": (ci, [C.c_char_p, ci, ci, C.POINTER(vp), C.c_char_p, cs_]),
    "fts_integrand_info": (ci, [vp] + [C.POINTER(ci)] * 3),
    "fts_integrand_free": (ci, [vp]),
    "fts_eval_batch": (ci, [vp, ci, vp, vp, ci, i64, vp]),
    "fts_integrate1d_cmpl_batch": (ci, [vp, vp, ci, vp, ci, i64, vp]),
    "fts_integrate1d_cmpl_batch_dev": (ci, [vp, vp, ci, vp, ci, i64, vp, vp]),
    "fts_integrate_batch": (ci, [vp, vp, ci, vp, vp, ci, vp, ci, i64, vp]),
    "fts_integrate_batch_dev": (ci, [vp, vp, ci, vp, vp, ci, vp, ci, i64, vp, vp]),
    "fts_adaptive_batch": (ci, [vp, vp, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci, i64, vp, vp, vp, vp]),
    "fts_adaptive_batch_dev": (ci, [vp, vp, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci, i64, vp, vp, vp, vp, vp]),
    "fts_adaptive_sharded_level_dev": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp]),
    "fts_adaptive_sharded_step_dev": (ci, [vp, vp, vp, vp, vp, ci, ci, vp, vp, vp]),
    "fts_adaptive_sharded_step_opts_dev": (ci, [vp, vp, vp, vp, vp, ci, ci, vp, vp, ci, ci, vp]),
    "fts_peer_mailbox_bytes": (ci, [ci, C.POINTER(cs_)]),
    "fts_peer_alloc": (ci, [cs_, C.POINTER(vp), vp]),
    "fts_peer_open": (ci, [vp, C.POINTER(vp)]),
    "fts_peer_close": (ci, [vp]),
    "fts_peer_free": (ci, [vp]),
    "fts_adaptive_sharded_fused_dev": (ci, [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, C.POINTER(vp), C.c_uint64, vp, vp]),
    "fts_peer_wait_ns": (ci, [ci, C.POINTER(C.c_uint64)]),
    "fts_peer_barrier_dev": (ci, [C.POINTER(vp), ci, ci, C.c_uint64, vp]),
    "fts_adaptive_sharded_fused_opts_dev": (ci, [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, C.POINTER(vp), C.c_uint64, vp, ci, vp]),
    "fts_integrand_cubin_size": (ci, [vp, C.POINTER(cs_), C.POINTER(ci)]),
    "fts_measure_fma_peak": (ci, [ci, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fts_launch_count": (i64, [ci]),
    "fts_evals_for_level": (i64, [ci, ci]),
    "fts_evals_fixed": (i64, [ci, ci]),
}
EXPORTED = tuple(_SIGS)
_missing = []
for _name, (_res, _args) in _SIGS.items():
    try:
        _f = getattr(lib, _name)
    except AttributeError:
        _missing.append(_name)
        continue
    _f.restype, _f.argtypes = _res, _args
MISSING = tuple(_missing)


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    lib.fts_last_error(buf, 2048)
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_DIMENSION_MISMATCH:
        raise DimensionMismatch(rc, msg)
    if rc == ERR_INVALID_ARGUMENT:
        raise ArgumentError(rc, msg)
    raise FtsError(rc, msg)
