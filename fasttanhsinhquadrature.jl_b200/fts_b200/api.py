"""Host-side mirror of the FastTanhSinhQuadrature.jl API over libfts_b200.so.

Same names, argument meaning, defaults and error behaviour as the reference's exported surface
(/root/reference/src/FastTanhSinhQuadrature.jl:44-51), with two additive changes that the GPU
boundary needs:

* the integrand `f` is an :class:`Integrand` (a builtin family or a CUDA-C body compiled with
  NVRTC) instead of a host closure, with per-integral parameters passed as ``params=``;
* every entry point is *batched*: bounds / params may be arrays, the result is then an array.

Julia is absent from this image, so this Python module is the executable stand-in for the Julia
shim in ``julia/FastTanhSinhQuadratureB200`` (same calls into the same C ABI).
"""
from __future__ import annotations

import ctypes as C
import math
import threading
import warnings
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import ArgumentError, DimensionMismatch, FtsError

Double64 = np.dtype([("hi", "<f8"), ("lo", "<f8")])

_DT_CODE = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64, Double64: L.DD64}
_CODE_DT = {v: k for k, v in _DT_CODE.items()}

FAMILIES = {
    "poly7": 0, "exp": 1, "gauss": 2, "runge": 3, "log1mx": 4, "log1px": 5, "invsqrt1mx2": 6, "sin2": 7,
    "poly6": 8, "invsqrtabs": 9,
    "c_invsqrt": 20, "c_log_bmx": 21, "c_log_xma": 22, "c_exp_inv_bmx": 23, "c_exp_plus": 24,
    "poly2_2d": 40, "exp_2d": 41, "x2py2": 42, "invsqrt_2d": 43, "invsqrtabs_2d": 44,
    "c2_invsqrt_bmx_ymc": 50, "c2_invsqrt_all": 51,
    "poly2_3d": 60, "exp_3d": 61, "x2y2z2": 62, "invsqrt_3d": 63, "rat_3d": 64, "log_3d": 65, "invsqrtabs_3d": 66,
}


class ConvergenceWarning(UserWarning):
    """mirror of the reference's `@warn "... reached max_levels ..."` (adaptive.jl:71-73)"""


def _dt(T) -> np.dtype:
    if isinstance(T, str) and T.lower() in ("dd", "dd64", "double64"):
        return Double64
    d = np.dtype(T)
    if d not in _DT_CODE:
        raise TypeError(f"unsupported element type {T!r}: use float32, float64 or Double64")
    return d


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def to_dd(values) -> np.ndarray:
    """Convert floats / decimal strings / mpmath numbers to a Double64 array (hi, lo)."""
    arr = np.atleast_1d(values)
    if arr.dtype == Double64:
        return arr.copy()
    out = np.zeros(arr.shape, Double64)
    if arr.dtype.kind in "fiu":
        out["hi"] = arr.astype(np.float64)
        return out
    import mpmath
    with mpmath.workprec(200):
        for idx, v in np.ndenumerate(arr):
            m = mpmath.mpf(v)
            hi = float(m)
            out[idx] = (hi, float(m - mpmath.mpf(hi)))
    return out


def dd_add(a, b) -> np.ndarray:
    """a + b of Double64 arrays on the host, the arithmetic of csrc/fts_dd.cuh operator+ (two TwoSums, two renormalisations)"""
    a, b = np.atleast_1d(a), np.atleast_1d(b)

    def two_sum(x, y):
        s = x + y
        bb = s - x
        return s, (x - (s - bb)) + (y - bb)

    def quick_two_sum(x, y):
        s = x + y
        return s, y - (s - x)

    sh, sl = two_sum(a["hi"], b["hi"])
    th, tl = two_sum(a["lo"], b["lo"])
    sh, sl = quick_two_sum(sh, sl + th)
    sh, sl = quick_two_sum(sh, sl + tl)
    out = np.zeros(np.shape(sh), Double64)
    out["hi"], out["lo"] = sh, sl
    return out


def dd_to_mpf(a):
    """Double64 array -> list of mpmath.mpf (exact hi+lo)."""
    import mpmath
    a = np.atleast_1d(a)
    with mpmath.workprec(200):
        return [mpmath.mpf(float(v["hi"])) + mpmath.mpf(float(v["lo"])) for v in a.reshape(-1)]


def _as_elems(v, dt: np.dtype) -> np.ndarray:
    """array of element type dt from user input (scalars, sequences, arrays)"""
    if dt == Double64:
        return np.ascontiguousarray(to_dd(v))
    return np.ascontiguousarray(np.asarray(v, dtype=dt))


def _scalar(v, dt: np.dtype) -> np.ndarray:
    return _as_elems(v, dt).reshape(-1)[:1].copy()


# ------------------------------------------------------------------------------- lifecycle ----
def init(device: int = -1) -> None:
    """Bind this process to CUDA device `device` (one process per GPU)."""
    L.check(L.lib.fts_init(device))


def device_info() -> dict:
    dev, sm, ma, mi = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    mem = C.c_size_t()
    L.check(L.lib.fts_device_info(C.byref(dev), C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
    return dict(device=dev.value, sm_count=sm.value, cc=(ma.value, mi.value), global_mem=mem.value)


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array in page-locked, device-mapped host memory (fts_host_alloc).  Per-integral operands
    and result arrays that live here are read and written IN PLACE by the kernels over PCIe
    (fts_integrate_batch / fts_adaptive_batch take no staging copy); freed with the array."""
    import weakref
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    ptr = C.c_void_p()
    L.check(L.lib.fts_host_alloc(C.byref(ptr), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(ptr.value)
    weakref.finalize(buf, L.lib.fts_host_free, C.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def launch_count(reset: bool = False) -> int:
    return int(L.lib.fts_launch_count(int(reset)))


def measure_fma_peak(T=np.float64, iters: int = 8192) -> dict:
    t, ms = C.c_double(), C.c_double()
    L.check(L.lib.fts_measure_fma_peak(_DT_CODE[_dt(T)], iters, C.byref(t), C.byref(ms)))
    return dict(tflops=t.value, ms=ms.value)


def evals_for_level(D: int, level) -> np.ndarray:
    """integrand evaluations of one integral that stopped at `level`: (2^(level+2)+1)^D"""
    m = (np.int64(2) << (np.asarray(level, np.int64) + 1)) + 1
    return m ** D


def adaptive_batch_dev(cache, f, *, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, value, err=None,
                       levels=None, flags=None, stream=None, order="reference", options=0) -> None:
    """Asynchronous launch on DEVICE buffers (raw pointers as ints, e.g. torch.Tensor.data_ptr()):
    fts_adaptive_batch_dev.  rtol/atol are host scalars of the cache's element type."""
    dt = cache.dtype
    r, a = _scalar(rtol, dt), _scalar(atol, dt)
    L.check(L.lib.fts_adaptive_batch_dev(cache.handle, f._h, _order(order), options, C.c_void_p(low), C.c_void_p(up), bounds_stride,
                                         C.c_void_p(params) if params else None, params_stride, _ptr(r), _ptr(a), max_levels, batch,
                                         C.c_void_p(value), C.c_void_p(err) if err else None, C.c_void_p(levels) if levels else None,
                                         C.c_void_p(flags) if flags else None, C.c_void_p(stream) if stream else None))


def integrate_batch_dev(x, w, h, f, *, low, up, bounds_stride, params, params_stride, batch, value, stream=None, order="reference") -> None:
    """Asynchronous fixed-grid launch on DEVICE buffers (raw pointers as ints): fts_integrate_batch_dev."""
    handle = _tables.get(x, w, h)
    L.check(L.lib.fts_integrate_batch_dev(handle, f._h, _order(order), C.c_void_p(low), C.c_void_p(up), bounds_stride,
                                          C.c_void_p(params) if params else None, params_stride, batch, C.c_void_p(value),
                                          C.c_void_p(stream) if stream else None))


def adaptive_batch_host(cache, f, *, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, value, err=None,
                        levels=None, flags=None, order="reference", options=0) -> None:
    """Synchronous call on HOST buffers (raw pointers; pinned memory makes the copies DMA):
    fts_adaptive_batch — the call a Julia `ccall` makes, PCIe copies included."""
    dt = cache.dtype
    r, a = _scalar(rtol, dt), _scalar(atol, dt)
    L.check(L.lib.fts_adaptive_batch(cache.handle, f._h, _order(order), options, C.c_void_p(low), C.c_void_p(up), bounds_stride,
                                     C.c_void_p(params) if params else None, params_stride, _ptr(r), _ptr(a), max_levels, batch,
                                     C.c_void_p(value), C.c_void_p(err) if err else None, C.c_void_p(levels) if levels else None,
                                     C.c_void_p(flags) if flags else None))


def sharded_level_dev(cache, f, *, low, up, params, level, rank, nranks, partial, state=None, stream=None) -> None:
    """fts_adaptive_sharded_level_dev on device pointers (ints)."""
    L.check(L.lib.fts_adaptive_sharded_level_dev(cache.handle, f._h, C.c_void_p(low), C.c_void_p(up), C.c_void_p(params) if params else None,
                                                 level, rank, nranks, C.c_void_p(partial), C.c_void_p(state) if state else None,
                                                 C.c_void_p(stream) if stream else None))


def sharded_step_dev(cache, *, low, up, rtol, atol, level, nranks, partials, state, stream=None, options=0, max_levels=0) -> None:
    """fts_adaptive_sharded_step_opts_dev on device pointers (ints)."""
    dt = cache.dtype
    r, a = _scalar(rtol, dt), _scalar(atol, dt)
    L.check(L.lib.fts_adaptive_sharded_step_opts_dev(cache.handle, C.c_void_p(low), C.c_void_p(up), _ptr(r), _ptr(a), level, nranks,
                                                     C.c_void_p(partials), C.c_void_p(state), options, max_levels, C.c_void_p(stream) if stream else None))


# ------------------------------------------------------------------------------- integrands ---
class Integrand:
    """A device integrand: builtin family (statically compiled sm_100a kernels) or an NVRTC module."""

    def __init__(self, handle, kind: int, nparams: int, name: str, builtin: bool):
        self._h, self.kind, self.nparams, self.name, self.is_builtin = handle, kind, nparams, name, builtin

    @property
    def dim(self) -> int:
        if self.kind > L.KIND_ND:
            return self.kind - L.KIND_ND
        return {L.KIND_1D: 1, L.KIND_1D_CMPL: 1, L.KIND_2D: 2, L.KIND_2D_CMPL: 2, L.KIND_3D: 3}[self.kind]

    def __del__(self):
        try:
            if self._h:
                L.lib.fts_integrand_free(self._h)
                self._h = None
        except Exception:
            pass

    def __repr__(self):
        return f"Integrand({self.name!r}, kind={self.kind}, nparams={self.nparams})"


_builtin_cache: dict = {}


def builtin(family) -> Integrand:
    """Builtin family by name (see FAMILIES) or numeric id (include/fts_b200.h fts_family)."""
    fid = FAMILIES[family] if isinstance(family, str) else int(family)
    if fid in _builtin_cache:
        return _builtin_cache[fid]
    h = C.c_void_p()
    L.check(L.lib.fts_integrand_builtin(fid, C.byref(h)))
    kind, npar, isb = C.c_int(), C.c_int(), C.c_int()
    L.check(L.lib.fts_integrand_info(h, C.byref(kind), C.byref(npar), C.byref(isb)))
    name = family if isinstance(family, str) else next((k for k, v in FAMILIES.items() if v == fid), str(fid))
    f = Integrand(h, kind.value, npar.value, name, True)
    _builtin_cache[fid] = f
    return f


_KIND_BY_NAME = {"1d": L.KIND_1D, "2d": L.KIND_2D, "3d": L.KIND_3D, "1d_cmpl": L.KIND_1D_CMPL, "cmpl": L.KIND_1D_CMPL,
                 "2d_cmpl": L.KIND_2D_CMPL}


def cuda_integrand(body: str, kind="1d", nparams: int = 0, name: str = "user", dim: Optional[int] = None) -> Integrand:
    """Compile a user integrand with NVRTC for sm_100a.

    `body` is the body of ``template<class T> __device__ T f(ARGS, const T* p)`` with ARGS =
    ``T x`` | ``T x, T y`` | ``T x, T y, T z`` | ``T x, T bmx, T xma`` |
    ``T x, T y, T bmx, T xma, T dmy, T ymc`` — e.g. ``"return exp(-p[0]*x*x);"``.
    kind="nd", dim=D (1 <= D <= 8): ARGS = ``const T* x`` (x[0] ... x[D-1]; ``ND`` is D) for integrateND.
    """
    if kind == "nd":
        if dim is None or not 1 <= int(dim) <= L.ND_MAX:
            raise ArgumentError(L.ERR_INVALID_ARGUMENT, f"cuda_integrand(kind='nd') needs dim = 1 ... {L.ND_MAX}")
        k = L.KIND_ND + int(dim)
    else:
        k = _KIND_BY_NAME[kind] if isinstance(kind, str) else int(kind)
    h = C.c_void_p()
    log = C.create_string_buffer(16384)
    rc = L.lib.fts_integrand_nvrtc(body.encode(), k, nparams, C.byref(h), log, len(log))
    if rc != L.OK:
        raise FtsError(rc, L.last_error() + "\n" + log.value.decode(errors="replace"))
    return Integrand(h, k, nparams, name, False)


def evaluate(f: "Integrand", x, params=None, T=np.float64) -> np.ndarray:
    """f(x) at the points x, evaluated on the device (fts_eval_batch; kind f(x), float32/float64)."""
    dt = _dt(T)
    xs = np.ascontiguousarray(np.asarray(x, dt).reshape(-1))
    p, pstride, _ = _params(f, params, dt)
    out = np.zeros(xs.shape[0], dt)
    code = L.F32 if dt == np.dtype(np.float32) else L.F64
    L.check(L.lib.fts_eval_batch(f._h, code, _ptr(xs), _ptr(p), pstride, xs.shape[0], _ptr(out)))
    return out


# ------------------------------------------------------------------------------- windows ------
def _window(T, which: int, D: int):
    dt = _dt(T)
    out = np.zeros(1, dt)
    L.check(L.lib.fts_window(_DT_CODE[dt], which, D, _ptr(out)))
    return out[0]


def tmax(T, D: int = 1):
    """tmax(T, D)   /root/reference/src/core.jl:228-230"""
    return _window(T, 0, D)


def t_x_max(T):
    """t_x_max(T)   core.jl:197-199"""
    return _window(T, 1, 1)


def t_w_max(T, D: int = 1):
    """t_w_max(T, D)   core.jl:207-221"""
    return _window(T, 2, D)


# ------------------------------------------------------------------------------- fixed grids --
def tanhsinh(T=np.float64, N: Optional[int] = None, D: int = 1):
    """tanhsinh(T, N; D) -> (x, w, h)   core.jl:256-267; tanhsinh(N) uses Float64 (core.jl:274)."""
    if N is None:
        T, N = np.float64, T
    dt = _dt(T)
    N = int(N)
    n = N // 2
    x = np.zeros(n, dt)
    w = np.zeros(n, dt)
    h = np.zeros(1, dt)
    L.check(L.lib.fts_tanhsinh(_DT_CODE[dt], N, D, _ptr(x), _ptr(w), _ptr(h)))
    return x, w, h[0]


def hopt(T=np.float64, N: int = 2, d: float = math.pi / 2):
    """hopt(T, N, d) = (2/N) W(2dN) (core.jl:289-291)"""
    dt = _dt(T)
    out = np.zeros(1, dt)
    L.check(L.lib.fts_hopt(_DT_CODE[dt], int(N), float(d), _ptr(out)))
    return out[0]


def hmax(T=np.float64, N: int = 2, D: int = 1):
    """hmax(T, N, D) = tmax(T, D) / (N div 2) (core.jl:293-302)"""
    dt = _dt(T)
    out = np.zeros(1, dt)
    L.check(L.lib.fts_hmax(_DT_CODE[dt], int(N), int(D), _ptr(out)))
    return out[0]


def tanhsinh_opt(T=np.float64, N: int = 2, d: float = math.pi / 2, D: int = 1):
    """tanhsinh_opt(T, N, d; D) (core.jl:276-287): nodes at t = h:h:tmax(T,D) with the
    optimal-spacing step h = hopt(T, N, d).  Returns (x, w, h)."""
    dt = _dt(T)
    h = np.zeros(1, dt)
    count = C.c_int64()
    L.check(L.lib.fts_tanhsinh_opt(_DT_CODE[dt], int(N), float(d), int(D), 0, None, None, _ptr(h), C.byref(count)))
    x, w = np.zeros(count.value, dt), np.zeros(count.value, dt)
    L.check(L.lib.fts_tanhsinh_opt(_DT_CODE[dt], int(N), float(d), int(D), count.value, _ptr(x), _ptr(w), _ptr(h), C.byref(count)))
    return x, w, h[0]


class _TableCache:
    """device copies of (x,w,h) triples handed to integrate*D, keyed by the host arrays' identity"""

    def __init__(self, cap: int = 16):
        self.cap, self.items, self.lock = cap, {}, threading.Lock()

    def get(self, x: np.ndarray, w: np.ndarray, h) -> C.c_void_p:
        key = (id(x), id(w), x.shape[0], x.dtype.str)
        with self.lock:
            hit = self.items.get(key)
            if hit is not None and hit[1] is x and hit[2] is w:
                return hit[0]
            if x.dtype != w.dtype or x.shape != w.shape or x.ndim != 1:
                raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, "x and w must be vectors of the same length and type")
            dt = x.dtype
            hh = _scalar(h, dt)
            handle = C.c_void_p()
            xc, wc = np.ascontiguousarray(x), np.ascontiguousarray(w)
            L.check(L.lib.fts_tables_fixed_upload(_DT_CODE[dt], xc.shape[0], _ptr(xc), _ptr(wc), _ptr(hh), C.byref(handle)))
            if len(self.items) >= self.cap:
                old = next(iter(self.items))
                L.lib.fts_tables_free(self.items.pop(old)[0])
            self.items[key] = (handle, x, w)
            return handle


_tables = _TableCache()


def _order(order) -> int:
    return {"reference": L.ORDER_REFERENCE, "fast": L.ORDER_FAST, "auto": L.ORDER_AUTO}[order] if isinstance(order, str) else int(order)


def _bounds(low, up, dim: int, dt: np.dtype, who: str):
    """-> (low_arr, up_arr, stride, batch or None)"""
    lo, hi = _as_elems(low, dt), _as_elems(up, dt)
    if dim == 1:
        lo, hi = lo.reshape(-1), hi.reshape(-1)
        if lo.shape != hi.shape:
            lo, hi = np.broadcast_arrays(lo, hi)
            lo, hi = np.ascontiguousarray(lo), np.ascontiguousarray(hi)
        scalar = np.ndim(low) == 0 and np.ndim(up) == 0
        return lo, hi, (0 if scalar else 1), (None if scalar else lo.shape[0])
    if lo.ndim == 1 and hi.ndim == 1:
        if lo.shape[0] != dim:
            raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"low must have length {dim}")
        if hi.shape[0] != dim:
            raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"up must have length {dim}")
        return lo, hi, 0, None
    if lo.ndim == 1:
        lo = np.ascontiguousarray(np.broadcast_to(lo, hi.shape))
    if hi.ndim == 1:
        hi = np.ascontiguousarray(np.broadcast_to(hi, lo.shape))
    if lo.shape != hi.shape or lo.shape[-1] != dim:
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"{who}: low/up must be [batch, {dim}]")
    return lo, hi, dim, lo.shape[0]


def _params(f: Integrand, params, dt: np.dtype):
    """-> (params_arr or None, stride, batch or None)"""
    if f.nparams == 0:
        return None, 0, None
    if params is None:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, f"integrand {f.name!r} needs params with {f.nparams} values per integral")
    p = _as_elems(params, dt)
    if p.ndim == 0:
        p = p.reshape(1)
    if p.ndim == 1 and f.nparams == 1 and p.shape[0] != 1:
        p = p.reshape(-1, 1)
    if p.ndim == 1:
        if p.shape[0] < f.nparams:
            raise ArgumentError(L.ERR_INVALID_ARGUMENT, f"integrand {f.name!r} needs {f.nparams} parameters")
        return p, 0, None
    if p.shape[1] < f.nparams:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, f"integrand {f.name!r} needs {f.nparams} parameters per integral")
    return np.ascontiguousarray(p), p.shape[1], p.shape[0]


def _batch_of(*sizes) -> tuple:
    known = [s for s in sizes if s is not None]
    if not known:
        return 1, True
    if any(s != known[0] for s in known):
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"inconsistent batch sizes {known}")
    return known[0], False


def _integrate(f: Integrand, dim: int, low, up, x, w, h, params, order, who: str):
    if not isinstance(f, Integrand):
        raise TypeError(f"{who}: f must be an fts_b200.Integrand (builtin(...) or cuda_integrand(...)); host closures cannot run on the GPU")
    if f.dim != dim or f.kind in (L.KIND_1D_CMPL, L.KIND_2D_CMPL):
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"{who} needs a {dim}D integrand, got {f!r}")
    dt = x.dtype
    lo, hi, bstride, nb = _bounds(low, up, dim, dt, who)
    p, pstride, npb = _params(f, params, dt)
    batch, scalar = _batch_of(nb, npb)
    out = np.zeros(batch, dt)
    handle = _tables.get(x, w, h)
    L.check(L.lib.fts_integrate_batch(handle, f._h, _order(order), _ptr(lo), _ptr(hi), bstride, _ptr(p), pstride, batch, _ptr(out)))
    return out[0] if scalar else out


def integrate1D(f, *args, params=None, order="reference"):
    """integrate1D(f, x, w, h) | integrate1D(f, low, up, x, w, h)   integrate1D.jl:74-113"""
    if len(args) == 3:
        x, w, h = args
        one = np.ones(1, x.dtype)[0] if x.dtype != Double64 else 1.0
        return _integrate(f, 1, -one, one, x, w, h, params, order, "integrate1D")
    low, up, x, w, h = args
    return _integrate(f, 1, low, up, x, w, h, params, order, "integrate1D")


def integrate1D_avx(f, *args, params=None):
    """integrate1D_avx   integrate1D.jl:124-157 (reassociated + FMA)"""
    return integrate1D(f, *args, params=params, order="fast")


def integrate1D_cmpl(T, f, N: int, params=None, order="reference"):
    """integrate1D_cmpl(T, f, N)   integrate1D.jl:49-65: integral of f(x, 1-x, 1+x) over [-1, 1] on the grid
    tanhsinh(T, N), complements formed by subtraction (unexported in the reference as well)."""
    if not isinstance(f, Integrand) or f.kind != L.KIND_1D_CMPL:
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, "integrate1D_cmpl needs an integrand of kind f(x, bmx, xma)")
    dt = _dt(T)
    x, w, h = tanhsinh(dt, N)
    handle = _tables.get(x, w, h)
    p, pstride, npb = _params(f, params, dt)
    batch = npb if npb else 1
    out = np.zeros(batch, dt)
    L.check(L.lib.fts_integrate1d_cmpl_batch(handle, f._h, _order(order), _ptr(p), pstride, batch, _ptr(out)))
    return out if npb else out[0]


def integrate2D(f, *args, params=None, order="reference"):
    """integrate2D(f, x, w, h) | integrate2D(f, low, up, x, w, h)   integrate2D.jl:80-157"""
    if len(args) == 3:
        x, w, h = args
        return _integrate(f, 2, [-1.0, -1.0], [1.0, 1.0], x, w, h, params, order, "integrate2D")
    low, up, x, w, h = args
    return _integrate(f, 2, low, up, x, w, h, params, order, "integrate2D")


def integrate2D_avx(f, low, up, x, w, h, params=None):
    """integrate2D_avx(f, low, up, x, w, h)   integrate2D.jl:164-202 (no unit-square method: App. B.7)"""
    return integrate2D(f, low, up, x, w, h, params=params, order="fast")


def integrate3D(f, *args, params=None, order="reference"):
    """integrate3D(f, x, w, h) | integrate3D(f, low, up, x, w, h)   integrate3D.jl:130-240"""
    if len(args) == 3:
        x, w, h = args
        return _integrate(f, 3, [-1.0] * 3, [1.0] * 3, x, w, h, params, order, "integrate3D")
    low, up, x, w, h = args
    return _integrate(f, 3, low, up, x, w, h, params, order, "integrate3D")


def integrateND(f, *args, params=None):
    """integrateND(f, x, w, h) | integrateND(f, low, up, x, w, h): the tensor-product node sum of integrate1D/2D/3D on a
    D-dimensional box, D = f.dim <= 8 (the generalisation the reference's paper names as its next step,
    JOSS_paper/paper.md:121; no reference implementation exists).  f = cuda_integrand(body, kind="nd", dim=D);
    float32 / float64; reassociated summation with compensated partials (the `_avx` flavour).  Build the grid with
    tanhsinh(T, N, D=D): the weight window depends on D (core.jl:207-221)."""
    if not isinstance(f, Integrand) or f.kind <= L.KIND_ND:
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, "integrateND needs an integrand of kind 'nd' (cuda_integrand(body, kind='nd', dim=D))")
    D = f.dim
    if len(args) == 3:
        x, w, h = args
        return _integrate(f, D, [-1.0] * D, [1.0] * D, x, w, h, params, "fast", "integrateND")
    low, up, x, w, h = args
    return _integrate(f, D, low, up, x, w, h, params, "fast", "integrateND")


def integrate3D_avx(f, *args, params=None):
    """integrate3D_avx   integrate3D.jl:141-143, 247-286"""
    return integrate3D(f, *args, params=params, order="fast")


# ------------------------------------------------------------------------------- caches -------
class AdaptiveCache:
    """_Adaptive1DCache / _AdaptiveTensorCache   caches.jl:31-39, 125-137.

    `xs/ws/cs` are per-level views (length(cache.xs) == max_levels) of the flat arrays that are
    uploaded to the device on first use; the per-level scratch vectors xp..zm of the reference
    do not exist (kernels keep them in shared memory)."""

    def __init__(self, dt, D, kind, max_levels, tm, ix, iw, ic, xs, ws, cs):
        self.dtype, self.D, self.kind, self.max_levels = dt, D, kind, max_levels
        self.tm, self.initial_x, self.initial_w, self.initial_c = tm, ix, iw, ic
        self._xs, self._ws, self._cs = xs, ws, cs
        off = (lambda l: (1 << l) - 2) if D == 0 else (lambda l: (2 << l) - 4)
        cnt = (lambda l: 1 << l) if D == 0 else (lambda l: 2 << l)
        self.xs = [xs[off(l):off(l) + cnt(l)] for l in range(1, max_levels + 1)]
        self.ws = [ws[off(l):off(l) + cnt(l)] for l in range(1, max_levels + 1)]
        self.cs = [cs[off(l):off(l) + cnt(l)] for l in range(1, max_levels + 1)] if cs is not None else []
        self._h = None
        self._lock = threading.Lock()

    @property
    def handle(self):
        with self._lock:
            if self._h is None:
                h = C.c_void_p()
                code = _DT_CODE[self.dtype]
                tm = np.array([self.tm], self.dtype)
                if self.D == 0:
                    L.check(L.lib.fts_cache1d_upload(code, self.kind, self.max_levels, _ptr(tm), _ptr(self.initial_x), _ptr(self.initial_w),
                                                     _ptr(self.initial_c), _ptr(self._xs), _ptr(self._ws), _ptr(self._cs), C.byref(h)))
                elif self.kind == 1:
                    L.check(L.lib.fts_cache_tensor_cmpl_upload(code, self.max_levels, _ptr(tm), _ptr(self.initial_x), _ptr(self.initial_w),
                                                               _ptr(self.initial_c), _ptr(self._xs), _ptr(self._ws), _ptr(self._cs), C.byref(h)))
                else:
                    L.check(L.lib.fts_cache_tensor_upload(code, self.D, self.max_levels, _ptr(tm), _ptr(self.initial_x), _ptr(self.initial_w),
                                                          _ptr(self._xs), _ptr(self._ws), C.byref(h)))
                self._h = h
            return self._h

    def __del__(self):
        try:
            if self._h:
                L.lib.fts_cache_free(self._h)
        except Exception:
            pass


_CACHES: dict = {}
_CACHES_LOCK = threading.RLock()  # caches.jl:50-51, 139-140


def _build_cache(dt, D, kind, max_levels) -> AdaptiveCache:
    if max_levels < 0:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.")
    code = _DT_CODE[dt]
    tm = np.zeros(1, dt); ix = np.zeros(2, dt); iw = np.zeros(2, dt); ic = np.zeros(2, dt)
    if D == 0:
        tot = (2 << max_levels) - 2
        xs = np.zeros(tot, dt); ws = np.zeros(tot, dt); cs = np.zeros(tot, dt)
        L.check(L.lib.fts_cache1d_generate(code, max_levels, kind, _ptr(tm), _ptr(ix), _ptr(iw), _ptr(ic), _ptr(xs), _ptr(ws), _ptr(cs)))
    elif kind == 1:
        tot = (4 << max_levels) - 4
        xs = np.zeros(tot, dt); ws = np.zeros(tot, dt); cs = np.zeros(tot, dt)
        L.check(L.lib.fts_cache_tensor_cmpl_generate(code, max_levels, _ptr(tm), _ptr(ix), _ptr(iw), _ptr(ic), _ptr(xs), _ptr(ws), _ptr(cs)))
    else:
        tot = (4 << max_levels) - 4
        xs = np.zeros(tot, dt); ws = np.zeros(tot, dt); cs = None
        L.check(L.lib.fts_cache_tensor_generate(code, D, max_levels, _ptr(tm), _ptr(ix), _ptr(iw), _ptr(xs), _ptr(ws)))
    return AdaptiveCache(dt, D, kind, max_levels, tm[0], ix, iw, ic, xs, ws, cs)


def _build_cache_on_device(dt, D, kind, max_levels) -> AdaptiveCache:
    """fts_cache_create: the level tables are computed by a kernel straight into device memory;
    the host copy behind `.xs/.ws/.cs` is a download of them."""
    if max_levels < 0:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.")
    code = _DT_CODE[dt]
    h = C.c_void_p()
    L.check(L.lib.fts_cache_create(code, D, kind, max_levels, C.byref(h)))
    tot = ((2 << max_levels) - 2) if D == 0 else ((4 << max_levels) - 4)
    tm = np.zeros(1, dt); ix = np.zeros(2, dt); iw = np.zeros(2, dt); ic = np.zeros(2, dt)
    xs = np.zeros(tot, dt); ws = np.zeros(tot, dt)
    cs = np.zeros(tot, dt) if (kind == 1 or D == 0) else None
    want_c = kind == 1
    L.check(L.lib.fts_cache_download(h, _ptr(tm), _ptr(ix), _ptr(iw), _ptr(ic) if want_c else None, _ptr(xs), _ptr(ws),
                                     _ptr(cs) if want_c else None))
    c = AdaptiveCache(dt, D, kind, max_levels, tm[0], ix, iw, ic, xs, ws, cs)
    c._h = h
    return c


def _cache(dt, D, kind, max_levels, on_device: bool = False) -> AdaptiveCache:
    key = (dt.str if dt != Double64 else "dd", D, kind, int(max_levels), bool(on_device))
    with _CACHES_LOCK:
        c = _CACHES.get(key)
        if c is None:
            c = _CACHES[key] = (_build_cache_on_device if on_device else _build_cache)(dt, D, kind, int(max_levels))
        return c


def adaptive_cache_1D(T=np.float64, max_levels: int = 16, complement: bool = False, on_device: bool = False) -> AdaptiveCache:
    """adaptive_cache_1D(T; max_levels=16, complement=false)   caches.jl:120-123 (memoised).
    on_device=True builds the tables with a kernel in HBM (fts_cache_create) instead of on the host."""
    return _cache(_dt(T), 0, 1 if complement else 0, max_levels, on_device)


def adaptive_cache_2D(T=np.float64, max_levels: int = 8, on_device: bool = False) -> AdaptiveCache:
    """adaptive_cache_2D(T; max_levels=8)   caches.jl:211-212"""
    return _cache(_dt(T), 2, 0, max_levels, on_device)


def adaptive_cache_3D(T=np.float64, max_levels: int = 5, on_device: bool = False) -> AdaptiveCache:
    """adaptive_cache_3D(T; max_levels=5)   caches.jl:219-220"""
    return _cache(_dt(T), 3, 0, max_levels, on_device)


def adaptive_cache_2D_cmpl(T=np.float64, max_levels: int = 8, on_device: bool = False) -> AdaptiveCache:
    """Tensor-complement cache (extension for BASELINE config C; window t_w_max(T, 2))."""
    return _cache(_dt(T), 2, 1, max_levels, on_device)


class DeviceTables:
    """Fixed tanh-sinh grid generated on the device (fts_tables_fixed_create): `x, w, h` are a
    download of it; `handle` goes to the *_dev launches."""

    def __init__(self, T=np.float64, N: int = 2, D: int = 1):
        dt = _dt(T)
        self.dtype = dt
        self.handle = C.c_void_p()
        L.check(L.lib.fts_tables_fixed_create(_DT_CODE[dt], int(N), int(D), C.byref(self.handle)))
        n = C.c_int()
        h = np.zeros(1, dt)
        L.check(L.lib.fts_tables_download(self.handle, C.byref(n), None, None, _ptr(h)))
        self.x, self.w = np.zeros(n.value, dt), np.zeros(n.value, dt)
        L.check(L.lib.fts_tables_download(self.handle, None, _ptr(self.x), _ptr(self.w), None))
        self.h = h[0]

    def __del__(self):
        try:
            L.lib.fts_tables_free(self.handle)
        except Exception:
            pass


def _require_cache_levels(cache: AdaptiveCache, max_levels: int) -> AdaptiveCache:
    """caches.jl:25-29"""
    if len(cache.xs) < max_levels:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT,
                            f"provided adaptive cache supports {len(cache.xs)} levels, but `max_levels={max_levels}` was requested.")
    return cache


# ------------------------------------------------------------------------------- tolerances ---
def _eps(dt: np.dtype) -> float:
    return 2.0 ** -104 if dt == Double64 else float(np.finfo(dt).eps)  # eps(Double64) in DoubleFloats.jl


def _resolve_tolerances(T, rtol=None, atol=0):
    """core.jl:32-44: rtol defaults to sqrt(eps(T)) iff atol == 0, else 0; both must be >= 0."""
    dt = _dt(T)
    atol_T = _scalar(atol, dt)
    atol_f = float(atol_T["hi"][0]) if dt == Double64 else float(atol_T[0])
    if rtol is not None:
        rtol_T = _scalar(rtol, dt)
    elif atol_f == 0:
        rtol_T = _scalar(math.sqrt(_eps(dt)), dt)
    else:
        rtol_T = _scalar(0.0, dt)
    rtol_f = float(rtol_T["hi"][0]) if dt == Double64 else float(rtol_T[0])
    if not rtol_f >= 0:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, "`rtol` must be nonnegative.")
    if not atol_f >= 0:
        raise ArgumentError(L.ERR_INVALID_ARGUMENT, "`atol` must be nonnegative.")
    return rtol_T, atol_T


# ------------------------------------------------------------------------------- adaptive -----
@dataclass
class AdaptiveResult:
    """value plus what the reference computes but does not return"""
    value: np.ndarray
    err: np.ndarray
    levels: np.ndarray
    flags: np.ndarray

    @property
    def converged(self):
        return (self.flags & L.FLAG_CONVERGED) != 0


def _tofloat(v):
    return float(v["hi"]) if isinstance(v, np.void) else float(v)


def _adaptive(name: str, T, f: Integrand, kind: int, low, up, *, rtol, atol, max_levels, warn, cache, params, order,
              options=0, full_output=False):
    if not isinstance(f, Integrand):
        raise TypeError(f"{name}: f must be an fts_b200.Integrand (builtin(...) or cuda_integrand(...)); host closures cannot run on the GPU")
    dt = _dt(T)
    if f.kind != kind:
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, f"{name} needs an integrand of kind {kind}, got {f!r}")
    rtol_T, atol_T = _resolve_tolerances(dt, rtol, atol)
    dim = f.dim
    if cache is None:
        if kind == L.KIND_1D:
            cache = adaptive_cache_1D(dt, max_levels=max_levels)
        elif kind == L.KIND_1D_CMPL:
            cache = adaptive_cache_1D(dt, max_levels=max_levels, complement=True)
        elif kind == L.KIND_2D:
            cache = adaptive_cache_2D(dt, max_levels=max_levels)
        elif kind == L.KIND_2D_CMPL:
            cache = adaptive_cache_2D_cmpl(dt, max_levels=max_levels)
        else:
            cache = adaptive_cache_3D(dt, max_levels=max_levels)
    else:
        _require_cache_levels(cache, max_levels)
    lo, hi, bstride, nb = _bounds(low, up, dim, dt, name)
    p, pstride, npb = _params(f, params, dt)
    batch, scalar = _batch_of(nb, npb)
    value = np.zeros(batch, dt); err = np.zeros(batch, dt)
    levels = np.zeros(batch, np.int32); flags = np.zeros(batch, np.int32)
    L.check(L.lib.fts_adaptive_batch(cache.handle, f._h, _order(order), options, _ptr(lo), _ptr(hi), bstride, _ptr(p), pstride,
                                     _ptr(rtol_T), _ptr(atol_T), max_levels, batch, _ptr(value), _ptr(err), _ptr(levels), _ptr(flags)))
    if warn and max_levels > 0:
        bad = np.nonzero(flags & L.FLAG_MAX_LEVELS)[0]
        if bad.size:
            i = int(bad[0])
            rt, at = _tofloat(rtol_T[0]), _tofloat(atol_T[0])
            target = max(at, rt * abs(_tofloat(value[i])))
            extra = f" ({bad.size} of {batch} integrals; first: index {i})" if batch > 1 else ""
            warnings.warn(f"{name} reached max_levels without meeting the requested tolerance.{extra} max_levels = {max_levels}, "
                          f"estimated_error = {_tofloat(err[i])}, target = {target}, value = {_tofloat(value[i])}, rtol = {rt}, atol = {at}",
                          ConvergenceWarning, stacklevel=3)
    if full_output:
        return AdaptiveResult(value, err, levels, flags)
    return value[0] if scalar else value


_ADAPT_KW = dict(rtol=None, atol=0, warn=True, cache=None, params=None, full_output=False)


# force_depth=True (additive, not in the reference): FTS_OPT_FORCE_DEPTH — every level up to max_levels is
# computed whatever the error estimate says.  The reference's own forced-depth idiom, rtol = atol = 0
# (benchmark/benchmarks.jl:95-110, runtests.jl:230), still stops when two levels agree to the last bit,
# which smooth integrands do at depth 6-7.
def adaptive_integrate_1D(T, f, a, b, *, rtol=None, atol=0, max_levels=16, warn=True, cache=None, params=None, order="reference", full_output=False, force_depth=False):
    """adaptive_integrate_1D(T, f, a, b; rtol, atol, max_levels=16, warn=true, cache)   adaptive.jl:28-85"""
    return _adaptive("adaptive_integrate_1D", T, f, L.KIND_1D, a, b, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order=order, full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_1D_avx(T, f, a, b, *, rtol=None, atol=0, max_levels=16, warn=True, cache=None, params=None, full_output=False, force_depth=False):
    """adaptive_integrate_1D_avx   adaptive.jl:94-143"""
    return _adaptive("adaptive_integrate_1D_avx", T, f, L.KIND_1D, a, b, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order="fast", full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_1D_cmpl(T, f, a, b, *, rtol=None, atol=0, max_levels=16, warn=True, cache=None, params=None, order="reference", full_output=False, force_depth=False):
    """adaptive_integrate_1D_cmpl   adaptive.jl:727-795; f(x, b-x, x-a)"""
    return _adaptive("adaptive_integrate_1D_cmpl", T, f, L.KIND_1D_CMPL, a, b, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn,
                     cache=cache, params=params, order=order, full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_1D_cmpl_avx(T, f, a, b, *, rtol=None, atol=0, max_levels=16, warn=True, cache=None, params=None, full_output=False, force_depth=False):
    """adaptive_integrate_1D_cmpl_avx   adaptive.jl:805-865"""
    return _adaptive("adaptive_integrate_1D_cmpl_avx", T, f, L.KIND_1D_CMPL, a, b, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn,
                     cache=cache, params=params, order="fast", full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_2D(T, f, low, up, *, rtol=None, atol=0, max_levels=8, warn=True, cache=None, params=None, order="reference", full_output=False, force_depth=False):
    """adaptive_integrate_2D   adaptive.jl:154-233"""
    return _adaptive("adaptive_integrate_2D", T, f, L.KIND_2D, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order=order, full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_2D_avx(T, f, low, up, *, rtol=None, atol=0, max_levels=8, warn=True, cache=None, params=None, full_output=False, force_depth=False):
    """adaptive_integrate_2D_avx   adaptive.jl:242-342"""
    return _adaptive("adaptive_integrate_2D_avx", T, f, L.KIND_2D, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order="fast", full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_2D_cmpl(T, f, low, up, *, rtol=None, atol=0, max_levels=8, warn=True, cache=None, params=None, order="reference", full_output=False, force_depth=False):
    """Tensor-complement adaptive 2D (extension, BASELINE config C): f(x, y, b-x, x-a, d-y, y-c)."""
    return _adaptive("adaptive_integrate_2D_cmpl", T, f, L.KIND_2D_CMPL, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn,
                     cache=cache, params=params, order=order, full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_3D(T, f, low, up, *, rtol=None, atol=0, max_levels=5, warn=True, cache=None, params=None, order="reference", full_output=False, force_depth=False):
    """adaptive_integrate_3D   adaptive.jl:352-480"""
    return _adaptive("adaptive_integrate_3D", T, f, L.KIND_3D, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order=order, full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


def adaptive_integrate_3D_avx(T, f, low, up, *, rtol=None, atol=0, max_levels=5, warn=True, cache=None, params=None, full_output=False, force_depth=False):
    """adaptive_integrate_3D_avx   adaptive.jl:489-715"""
    return _adaptive("adaptive_integrate_3D_avx", T, f, L.KIND_3D, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=warn, cache=cache,
                     params=params, order="fast", full_output=full_output, options=L.OPT_FORCE_DEPTH if force_depth else 0)


# ------------------------------------------------------------------------------- quad ---------
def _float_type(*vals):
    """promote(float(low), float(up)) (quad.jl:18): the result type is the bounds' float type"""
    dts = []
    for v in vals:
        a = np.asarray(v)
        if a.dtype == Double64:
            return Double64
        dts.append(a.dtype if a.dtype.kind == "f" else np.dtype(np.float64))
    r = np.result_type(*dts)
    return np.dtype(np.float64) if r not in (np.dtype(np.float32), np.dtype(np.float64)) else r


def quad(f, low=None, up=None, *, rtol=None, atol=0, max_levels=None, cache=None, params=None, order="reference", full_output=False):
    """quad(f, [low, up]; rtol, atol, max_levels, cache)   quad.jl:36-162.

    1D for scalar (or [batch]) bounds with a 1D integrand, 2D/3D for bounds of length 2/3
    (or [batch, D]); zero-width -> 0, flipped bounds -> negated (applied on the device)."""
    if low is None and up is None:
        low, up = -1.0, 1.0  # quad.jl:64-65
    if not isinstance(f, Integrand):
        raise TypeError("quad: f must be an fts_b200.Integrand")
    dim = f.dim
    if dim > 1:
        n_lo, n_up = np.shape(low)[-1] if np.ndim(low) else 0, np.shape(up)[-1] if np.ndim(up) else 0
        if n_lo != n_up:
            raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, "low and up must have the same length")  # quad.jl:70
        if n_lo > 3:
            raise FtsError(L.ERR_UNSUPPORTED, "Higher than 3D integration is not yet supported.")  # quad.jl:88
    T = _float_type(low, up)
    ml = max_levels if max_levels is not None else {1: 16, 2: 8, 3: 5}[dim]
    name = {L.KIND_1D: "adaptive_integrate_1D", L.KIND_2D: "adaptive_integrate_2D", L.KIND_3D: "adaptive_integrate_3D"}.get(f.kind)
    if name is None:
        raise DimensionMismatch(L.ERR_DIMENSION_MISMATCH, "quad needs a 1D/2D/3D integrand; use quad_cmpl for complement integrands")
    return _adaptive(name, T, f, f.kind, low, up, rtol=rtol, atol=atol, max_levels=ml, warn=True, cache=cache, params=params,
                     order=order, options=L.OPT_QUAD_EDGE_RULES, full_output=full_output)


def quad_cmpl(f, low=None, up=None, *, rtol=None, atol=0, max_levels=16, cache=None, params=None, order="reference", full_output=False):
    """quad_cmpl(f, [low, up]; ...)   quad.jl:298-318; f(x, b-x, x-a)"""
    if low is None and up is None:
        low, up = -1.0, 1.0
    if not isinstance(f, Integrand):
        raise TypeError("quad_cmpl: f must be an fts_b200.Integrand")
    T = _float_type(low, up)
    if f.kind == L.KIND_2D_CMPL:
        return _adaptive("adaptive_integrate_2D_cmpl", T, f, f.kind, low, up, rtol=rtol, atol=atol, max_levels=min(max_levels, 8) if max_levels == 16 else max_levels,
                         warn=True, cache=cache, params=params, order=order, options=L.OPT_QUAD_EDGE_RULES, full_output=full_output)
    return _adaptive("adaptive_integrate_1D_cmpl", T, f, L.KIND_1D_CMPL, low, up, rtol=rtol, atol=atol, max_levels=max_levels, warn=True,
                     cache=cache, params=params, order=order, options=L.OPT_QUAD_EDGE_RULES, full_output=full_output)


def quad_split(f, c, low=None, up=None, *, rtol=None, atol=0, max_levels=None, cache=None, params=None, order="reference"):
    """quad_split(f, c, [low, up]; ...)   quad.jl:164-281.

    The 2 / 4 / 8 sub-boxes of every integral are fanned out as ONE device batch (tolerances
    divided by 2 / 4 / 8 as in the reference) and summed on the host in the reference's order."""
    if not isinstance(f, Integrand):
        raise TypeError("quad_split: f must be an fts_b200.Integrand")
    dim = f.dim
    if low is None and up is None:
        low, up = (-1.0, 1.0) if dim == 1 else ([-1.0] * dim, [1.0] * dim)
    T = _float_type(c, low, up)
    dt = _dt(T)
    ml = max_levels if max_levels is not None else {1: 16, 2: 8, 3: 5}[dim]
    nsub = 1 << dim
    def div_tol(v):  # Double64 tolerances may be decimal strings / mpmath numbers
        if isinstance(v, str) or type(v).__module__.split(".")[0] == "mpmath":
            import mpmath
            with mpmath.workprec(200):
                return mpmath.mpf(v) / nsub
        return v / nsub

    sub_rtol = None if rtol is None else div_tol(rtol)
    sub_atol = div_tol(atol)
    if dt == Double64:
        cc, lo, hi = (to_dd(v).reshape(np.shape(v) if np.ndim(v) else ()) for v in (c, low, up))
    else:
        cc = np.asarray(c, dt); lo = np.asarray(low, dt); hi = np.asarray(up, dt)
    if dim == 1:
        cc, lo, hi = np.broadcast_arrays(cc.reshape(-1), lo.reshape(-1), hi.reshape(-1))
        scalar = np.ndim(c) == 0 and np.ndim(low) == 0 and np.ndim(up) == 0
        nb = cc.shape[0]
        sub_lo = np.stack([lo, cc], axis=1).reshape(-1)   # [low,c], [c,up]   quad.jl:172-173
        sub_hi = np.stack([cc, hi], axis=1).reshape(-1)
    else:
        cc, lo, hi = (np.atleast_2d(v) for v in (cc, lo, hi))
        nb = max(cc.shape[0], lo.shape[0], hi.shape[0])
        cc, lo, hi = (np.broadcast_to(v, (nb, dim)) for v in (cc, lo, hi))
        scalar = np.ndim(c) == 1 and np.ndim(low) == 1 and np.ndim(up) == 1
        edges = np.stack([lo, cc, hi], axis=1)  # [nb, 3, dim]
        sub_lo = np.zeros((nb, nsub, dim), dt); sub_hi = np.zeros((nb, nsub, dim), dt)
        if dim == 2:
            # q1 (x1,y1)-(cx,cy), q2 (cx,y1)-(x2,cy), q3 (x1,cy)-(cx,y2), q4 (cx,cy)-(x2,y2)   quad.jl:212-215
            order_idx = [(0, 0), (1, 0), (0, 1), (1, 1)]
            for s, (ix_, iy_) in enumerate(order_idx):
                sub_lo[:, s, 0] = edges[:, ix_, 0]; sub_hi[:, s, 0] = edges[:, ix_ + 1, 0]
                sub_lo[:, s, 1] = edges[:, iy_, 1]; sub_hi[:, s, 1] = edges[:, iy_ + 1, 1]
        else:
            s = 0
            for i in range(2):          # for i in 1:2, j in 1:2, k in 1:2 (k fastest)   quad.jl:262-264
                for j in range(2):
                    for k in range(2):
                        sub_lo[:, s, 0] = edges[:, i, 0]; sub_hi[:, s, 0] = edges[:, i + 1, 0]
                        sub_lo[:, s, 1] = edges[:, j, 1]; sub_hi[:, s, 1] = edges[:, j + 1, 1]
                        sub_lo[:, s, 2] = edges[:, k, 2]; sub_hi[:, s, 2] = edges[:, k + 1, 2]
                        s += 1
        sub_lo = sub_lo.reshape(nb * nsub, dim); sub_hi = sub_hi.reshape(nb * nsub, dim)
    p = params
    if p is not None and f.nparams:
        pa = np.asarray(p, dt)
        if pa.ndim == 2 or (pa.ndim == 1 and f.nparams == 1 and pa.shape[0] == nb and nb > 1):
            pa = pa.reshape(nb, -1)
            p = np.repeat(pa, nsub, axis=0)
    vals = quad(f, sub_lo, sub_hi, rtol=sub_rtol, atol=sub_atol, max_levels=ml, cache=cache, params=p, order=order)
    vals = np.asarray(vals).reshape(nb, nsub)
    total = vals[:, 0].copy()
    for s in range(1, nsub):
        total = dd_add(total, vals[:, s]) if dt == Double64 else total + vals[:, s]
    return total[0] if scalar else total
