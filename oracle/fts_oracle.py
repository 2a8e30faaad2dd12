"""ctypes front-end of the CPU oracle (oracle/fts_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs, never by the product package.

The oracle restates the *scalar* kernels of FastTanhSinhQuadrature.jl v2.1.0
(/root/reference/src/{core,caches,integrate1D,integrate2D,integrate3D,adaptive}.jl); each C
function cites the reference lines it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

# family ids (mirror of include/fts_b200.h; checked against the header by tests/test_abi.py)
FAMILIES = {
    "F1_POLY7": 0, "F1_EXP": 1, "F1_GAUSS": 2, "F1_RUNGE": 3, "F1_LOG1MX": 4, "F1_LOG1PX": 5,
    "F1_INVSQRT1MX2": 6, "F1_SIN2": 7, "F1_POLY6": 8, "F1_INVSQRTABS": 9,
    "C1_INVSQRT": 20, "C1_LOG_BMX": 21, "C1_LOG_XMA": 22, "C1_EXP_INV_BMX": 23, "C1_EXP_PLUS": 24,
    "F2_POLY2": 40, "F2_EXP": 41, "F2_X2PY2": 42, "F2_INVSQRT": 43, "F2_INVSQRTABS": 44,
    "C2_INVSQRT_BMX_YMC": 50, "C2_INVSQRT_ALL": 51,
    "F3_POLY2": 60, "F3_EXP": 61, "F3_X2Y2Z2": 62, "F3_INVSQRT": 63, "F3_RAT": 64, "F3_LOG": 65,
    "F3_INVSQRTABS": 66,
}


def native_tag() -> str:
    """Tag of this host's CPU (hash of its model name and ISA flags): -march=native builds are named
    after it, so a library built on another machine (the built files travel with the repository
    snapshot) is never loaded here."""
    import hashlib
    ident = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith(("model name", "flags")):
                    ident += line
                if line.strip() == "" and ident:
                    break
    except OSError:
        ident = "unknown"
    return hashlib.sha1(ident.encode()).hexdigest()[:10]


def build(native: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc, a few seconds)."""
    target = ["native", f"NATIVE_TAG={native_tag()}"] if native else []
    subprocess.run(["make", "-C", _HERE, "-s"] + target, check=True)


def _np_t(dtype: str):
    return np.float32 if dtype.startswith("f32") else np.float64


def _c_t(dtype: str):
    return C.c_float if dtype.startswith("f32") else C.c_double


@dataclass
class Result:
    value: float
    err: float
    level: int
    converged: bool
    evals: int


class _Res32(C.Structure):
    _fields_ = [("value", C.c_float), ("err", C.c_float), ("level", C.c_int), ("converged", C.c_int),
                ("evals", C.c_long)]


class _Res64(C.Structure):
    _fields_ = [("value", C.c_double), ("err", C.c_double), ("level", C.c_int), ("converged", C.c_int),
                ("evals", C.c_long)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """variant: 'strict' (parity checker), 'fast' (-O3, x86-64-v3), 'native' (-O3 -march=native)."""

    def __init__(self, variant: str = "strict"):
        name = {"strict": "libfts_oracle.so", "fast": "libfts_oracle_fast.so",
                "native": f"libfts_oracle_native_{native_tag()}.so"}[variant]
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build(native=(variant == "native"))
        self.lib = C.CDLL(path)
        self.variant = variant
        L = self.lib
        for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
            for fn in ("ordinate", "weight", "ordinate_complement", "inv_ordinate"):
                f = getattr(L, f"fts_or_{fn}_{sfx}")
                f.restype, f.argtypes = ct, [ct]
            getattr(L, f"fts_or_t_x_max_{sfx}").restype = ct
            getattr(L, f"fts_or_t_x_max_{sfx}").argtypes = []
            for fn in ("t_w_max", "tmax"):
                f = getattr(L, f"fts_or_{fn}_{sfx}")
                f.restype, f.argtypes = ct, [C.c_int]
        for sfx, ct in (("f32", C.c_float), ("f32x", C.c_float), ("f64", C.c_double), ("f64x", C.c_double)):
            vp = C.c_void_p
            f = getattr(L, f"fts_or_integrate1d_{sfx}")
            f.restype = ct
            f.argtypes = [C.c_int, vp, vp, ct, ct, vp, vp, C.c_int, ct]
            f = getattr(L, f"fts_or_integrate1d_cmpl_{sfx}")
            f.restype = ct
            f.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int, ct]
            for d in ("2d", "3d"):
                f = getattr(L, f"fts_or_integrate{d}_{sfx}")
                f.restype = ct
                f.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, ct]
            f = getattr(L, f"fts_or_adaptive1d_{sfx}")
            f.restype = None
            f.argtypes = [C.c_int, vp, vp, ct, ct, ct, ct, C.c_int, ct, vp, vp, vp, vp, vp]
            f = getattr(L, f"fts_or_adaptive1d_cmpl_{sfx}")
            f.restype = None
            f.argtypes = [C.c_int, vp, vp, ct, ct, ct, ct, C.c_int, ct, vp, vp, vp, vp, vp, vp, vp]
            for d in ("2d", "3d"):
                f = getattr(L, f"fts_or_adaptive{d}_{sfx}")
                f.restype = None
                f.argtypes = [C.c_int, vp, vp, vp, vp, ct, ct, C.c_int, ct, vp, vp, vp, vp, vp]
            f = getattr(L, f"fts_or_adaptive2d_cmpl_{sfx}")
            f.restype = None
            f.argtypes = [C.c_int, vp, vp, vp, ct, ct, C.c_int, ct, vp, vp, vp, vp, vp, vp, vp]
        L.fts_or_num_threads.restype = C.c_int
        L.fts_or_adaptive3d_omp_f64x.restype = None
        L.fts_or_adaptive3d_omp_f64x.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fts_or_adaptive3d_level_sum_omp_f64.restype = C.c_double
        L.fts_or_adaptive3d_level_sum_omp_f64.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                          C.c_void_p, C.c_void_p, C.c_long, C.c_int]

    # ------------------------------------------------------------------ primitives ----------
    def prim(self, name: str, dtype: str, *args):
        return getattr(self.lib, f"fts_or_{name}_{dtype}")(*args)

    def tmax(self, dtype: str, D: int = 1):
        return self.prim("tmax", dtype, D)

    def t_x_max(self, dtype: str):
        return self.prim("t_x_max", dtype)

    def t_w_max(self, dtype: str, D: int = 1):
        return self.prim("t_w_max", dtype, D)

    def num_threads(self) -> int:
        return self.lib.fts_or_num_threads()

    # ------------------------------------------------------------------ tables --------------
    def tanhsinh(self, dtype: str, N: int, D: int = 1):
        """(x, w, h) = tanhsinh(T, N; D)   src/core.jl:256-267"""
        if dtype == "dd":
            n = N // 2
            x = np.zeros(n, dtype="V16"); w = np.zeros(n, dtype="V16"); h = np.zeros(1, dtype="V16")
            self.lib.fts_or_tanhsinh_dd(C.c_int(N), C.c_int(D), _ptr(x), _ptr(w), _ptr(h))
            return x, w, h
        nt = _np_t(dtype)
        n = N // 2
        x = np.zeros(n, nt); w = np.zeros(n, nt); h = np.zeros(1, nt)
        getattr(self.lib, f"fts_or_tanhsinh_{dtype}")(C.c_int(N), C.c_int(D), _ptr(x), _ptr(w), _ptr(h))
        return x, w, nt(h[0])

    def cache1d(self, dtype: str, max_levels: int, complement: bool = False) -> dict:
        """_build_adaptive_1d_cache   src/caches.jl:59-91 (flat: level l at offset 2^l-2)"""
        nt = "V16" if dtype == "dd" else _np_t(dtype)
        tot = (2 << max_levels) - 2
        c = dict(tm=np.zeros(1, nt), ix=np.zeros(2, nt), iw=np.zeros(2, nt), ic=np.zeros(2, nt),
                 xs=np.zeros(tot, nt), ws=np.zeros(tot, nt), cs=np.zeros(tot, nt),
                 max_levels=max_levels, complement=bool(complement), dtype=dtype, D=0)
        getattr(self.lib, f"fts_or_cache1d_{dtype}")(
            C.c_int(max_levels), C.c_int(int(complement)), _ptr(c["tm"]), _ptr(c["ix"]), _ptr(c["iw"]),
            _ptr(c["ic"]), _ptr(c["xs"]), _ptr(c["ws"]), _ptr(c["cs"]))
        return c

    def cache_tensor(self, dtype: str, D: int, max_levels: int) -> dict:
        """_build_adaptive_tensor_cache   src/caches.jl:148-188 (flat: level l at offset 2^(l+1)-4)"""
        nt = "V16" if dtype == "dd" else _np_t(dtype)
        tot = (4 << max_levels) - 4
        c = dict(tm=np.zeros(1, nt), ix=np.zeros(2, nt), iw=np.zeros(2, nt),
                 xs=np.zeros(tot, nt), ws=np.zeros(tot, nt), max_levels=max_levels, dtype=dtype, D=D)
        getattr(self.lib, f"fts_or_cache_tensor_{dtype}")(
            C.c_int(D), C.c_int(max_levels), _ptr(c["tm"]), _ptr(c["ix"]), _ptr(c["iw"]),
            _ptr(c["xs"]), _ptr(c["ws"]))
        return c

    def cache_tensor_cmpl(self, dtype: str, max_levels: int) -> dict:
        """Tensor level table on the complement window t_w_max(T,2) with complements (extension
        used by BASELINE config C; nodes t=i*h_l as caches.jl:173-177, c as caches.jl:84).
        The window uses D=2 (core.jl:207-221) so that products of two weights / two complements
        stay above floatmin at the corners (with the 1D window w_i*w_j*f hits 0*Inf = NaN)."""
        nt = _np_t(dtype)
        sfx = dtype.rstrip("x")
        tm = self.t_w_max(sfx, 2)
        tot = (4 << max_levels) - 4
        xs = np.zeros(tot, nt); ws = np.zeros(tot, nt); cs = np.zeros(tot, nt)
        half = nt(0.5)
        h = nt(tm) * half
        o = lambda t: self.prim("ordinate", sfx, t)
        w_ = lambda t: self.prim("weight", sfx, t)
        c_ = lambda t: self.prim("ordinate_complement", sfx, t)
        h1, h2 = h, h + h
        ix = np.array([o(h1), o(h2)], nt); iw = np.array([w_(h1), w_(h2)], nt); ic = np.array([c_(h1), c_(h2)], nt)
        for level in range(1, max_levels + 1):
            h = h * half
            n = 2 << level
            off = n - 4
            for i in range(1, n + 1):
                t = nt(i) * h
                xs[off + i - 1] = o(t); ws[off + i - 1] = w_(t); cs[off + i - 1] = c_(t)
        return dict(tm=np.array([tm], nt), ix=ix, iw=iw, ic=ic, xs=xs, ws=ws, cs=cs,
                    max_levels=max_levels, dtype=dtype, D=2, complement=True)

    # ------------------------------------------------------------------ fixed grids ---------
    def _p(self, dtype, p):
        nt = _np_t(dtype)
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(p if p is not None else [0.0], dtype=nt)))
        if a.size < 8:
            a = np.concatenate([a, np.zeros(8 - a.size, nt)])
        return a

    def integrate1d(self, dtype, fam, p, low, up, x, w, h, cb=None):
        """integrate1D(f, low, up, x, w, h)   src/integrate1D.jl:90-102"""
        ct = _c_t(dtype); pa = self._p(dtype, p)
        return getattr(self.lib, f"fts_or_integrate1d_{dtype}")(
            fam, _ptr(pa), cb, ct(low), ct(up), _ptr(x), _ptr(w), len(x), ct(h))

    def integrate1d_cmpl(self, dtype, fam, p, x, w, h, cb=None):
        """integrate1D_cmpl(T, f, N) on the grid (x, w, h) = tanhsinh(T, N)   src/integrate1D.jl:49-65"""
        ct = _c_t(dtype); pa = self._p(dtype, p)
        return getattr(self.lib, f"fts_or_integrate1d_cmpl_{dtype}")(fam, _ptr(pa), cb, _ptr(x), _ptr(w), len(x), ct(h))

    def integrate_nd(self, dtype, D, fam, p, low, up, x, w, h, cb=None):
        """integrate2D / integrate3D   src/integrate2D.jl:91-124, src/integrate3D.jl:150-207"""
        nt = _np_t(dtype); ct = _c_t(dtype); pa = self._p(dtype, p)
        lo = np.ascontiguousarray(low, nt); hi = np.ascontiguousarray(up, nt)
        return getattr(self.lib, f"fts_or_integrate{D}d_{dtype}")(
            fam, _ptr(pa), cb, _ptr(lo), _ptr(hi), _ptr(x), _ptr(w), len(x), ct(h))

    @staticmethod
    def integrate_tensor_generic(fn, low, up, x, w, h):
        """The tensor-product rule of integrate2D / integrate3D (src/integrate2D.jl:91-124, src/integrate3D.jl:150-207)
        written for any dimension D = len(low): sum over the (2n+1)^D points (centre weight pi/2, nodes +-x_k) of
        prod_d w f(x0 + dx * xi), times prod_d dx_d * h^D.  The reference has no D > 3 (JOSS_paper/paper.md:121 names it
        as future work): this numpy statement is pinned by agreeing with integrate_nd for D = 2, 3.  `fn` takes an array
        [D, npoints] and returns [npoints]; float64, math.fsum (exact) summation.  Test infrastructure only."""
        import math
        lo = np.asarray(low, np.float64); hi = np.asarray(up, np.float64)
        D = lo.shape[0]
        xs = np.concatenate([[0.0], np.stack([np.asarray(x, np.float64), -np.asarray(x, np.float64)], 1).reshape(-1)])
        ws = np.concatenate([[math.pi / 2], np.repeat(np.asarray(w, np.float64), 2)])
        dx, x0 = 0.5 * (hi - lo), 0.5 * (hi + lo)
        grids = np.meshgrid(*[np.arange(xs.shape[0])] * D, indexing="ij")
        idx = np.stack([g.reshape(-1) for g in grids])
        pts = x0[:, None] + dx[:, None] * xs[idx]
        wt = np.prod(ws[idx], axis=0)
        return float(np.prod(dx) * float(h) ** D * math.fsum(wt * fn(pts)))

    # ------------------------------------------------------------------ adaptive ------------
    def _res(self, dtype):
        return _Res32() if dtype.startswith("f32") else _Res64()

    @staticmethod
    def _out(r) -> Result:
        return Result(r.value, r.err, r.level, bool(r.converged), r.evals)

    def adaptive1d(self, dtype, fam, p, a, b, rtol, atol, max_levels, cache, cb=None) -> Result:
        """adaptive_integrate_1D typed core   src/adaptive.jl:28-75"""
        ct = _c_t(dtype); pa = self._p(dtype, p); r = self._res(dtype)
        getattr(self.lib, f"fts_or_adaptive1d_{dtype}")(
            fam, _ptr(pa), cb, ct(a), ct(b), ct(rtol), ct(atol), max_levels, ct(cache["tm"][0]),
            _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["xs"]), _ptr(cache["ws"]), C.byref(r))
        return self._out(r)

    def adaptive1d_cmpl(self, dtype, fam, p, a, b, rtol, atol, max_levels, cache, cb=None) -> Result:
        """adaptive_integrate_1D_cmpl typed core   src/adaptive.jl:727-785"""
        ct = _c_t(dtype); pa = self._p(dtype, p); r = self._res(dtype)
        getattr(self.lib, f"fts_or_adaptive1d_cmpl_{dtype}")(
            fam, _ptr(pa), cb, ct(a), ct(b), ct(rtol), ct(atol), max_levels, ct(cache["tm"][0]),
            _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["ic"]), _ptr(cache["xs"]), _ptr(cache["ws"]),
            _ptr(cache["cs"]), C.byref(r))
        return self._out(r)

    def adaptive_nd(self, dtype, D, fam, p, low, up, rtol, atol, max_levels, cache, cb=None) -> Result:
        """adaptive_integrate_2D / _3D typed cores   src/adaptive.jl:154-224, 352-471"""
        nt = _np_t(dtype); ct = _c_t(dtype); pa = self._p(dtype, p); r = self._res(dtype)
        lo = np.ascontiguousarray(low, nt); hi = np.ascontiguousarray(up, nt)
        getattr(self.lib, f"fts_or_adaptive{D}d_{dtype}")(
            fam, _ptr(pa), cb, _ptr(lo), _ptr(hi), ct(rtol), ct(atol), max_levels, ct(cache["tm"][0]),
            _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["xs"]), _ptr(cache["ws"]), C.byref(r))
        return self._out(r)

    def adaptive3d_omp_f64x(self, fam, p, low, up, rtol, atol, max_levels, cache, nthreads: int = 0) -> Result:
        """adaptive_integrate_3D (src/adaptive.jl:352-471), exactly summed (__float128), every level
        split over OpenMP threads along the outer index: config D depths (L = 7, 8) in seconds."""
        pa = self._p("f64", p); r = _Res64()
        lo = np.ascontiguousarray(low, np.float64); hi = np.ascontiguousarray(up, np.float64)
        self.lib.fts_or_adaptive3d_omp_f64x(fam, _ptr(pa), _ptr(lo), _ptr(hi), rtol, atol, max_levels, float(cache["tm"][0]),
                                            _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["xs"]), _ptr(cache["ws"]), nthreads, C.byref(r))
        return self._out(r)

    def adaptive2d_cmpl(self, dtype, fam, p, low, up, rtol, atol, max_levels, cache) -> Result:
        nt = _np_t(dtype); ct = _c_t(dtype); pa = self._p(dtype, p); r = self._res(dtype)
        lo = np.ascontiguousarray(low, nt); hi = np.ascontiguousarray(up, nt)
        getattr(self.lib, f"fts_or_adaptive2d_cmpl_{dtype}")(
            fam, _ptr(pa), _ptr(lo), _ptr(hi), ct(rtol), ct(atol), max_levels, ct(cache["tm"][0]),
            _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["ic"]), _ptr(cache["xs"]), _ptr(cache["ws"]),
            _ptr(cache["cs"]), C.byref(r))
        return self._out(r)

    # ------------------------------------------------------------------ batches -------------
    def batch_adaptive1d(self, dtype, fam, params, low, up, rtol, atol, max_levels, cache,
                         nthreads: int = 0, cmpl: bool = False):
        """Loop of adaptive_integrate_1D[_cmpl] over a batch (README.md:96-106) under OpenMP."""
        nt = _np_t(dtype); ct = _c_t(dtype)
        low = np.ascontiguousarray(low, nt).reshape(-1); up = np.ascontiguousarray(up, nt).reshape(-1)
        params = np.ascontiguousarray(params, nt)
        batch = max(low.size, params.shape[0] if params.ndim else 1)
        pstride = 0 if params.ndim < 2 and params.size <= 1 else (params.shape[1] if params.ndim == 2 else 1)
        if params.ndim < 2 and params.size > 1:
            batch = max(batch, params.size)
        pbuf = params.reshape(-1)
        if pbuf.size == 0:
            pbuf = np.zeros(8, nt)
        if pstride == 0 and pbuf.size < 8:
            pbuf = np.concatenate([pbuf, np.zeros(8 - pbuf.size, nt)])
        bstride = 0 if low.size == 1 else 1
        value = np.zeros(batch, nt); err = np.zeros(batch, nt)
        level = np.zeros(batch, np.int32); conv = np.zeros(batch, np.int32); evals = np.zeros(batch, np.int64)
        if cmpl:
            getattr(self.lib, f"fts_or_batch_adaptive1d_cmpl_{dtype}")(
                fam, _ptr(pbuf), pstride, _ptr(low), _ptr(up), bstride, ct(rtol), ct(atol), max_levels,
                ct(cache["tm"][0]), _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["ic"]), _ptr(cache["xs"]),
                _ptr(cache["ws"]), _ptr(cache["cs"]), C.c_long(batch), _ptr(value), _ptr(err), _ptr(level),
                _ptr(conv), _ptr(evals), nthreads)
        else:
            getattr(self.lib, f"fts_or_batch_adaptive1d_{dtype}")(
                fam, _ptr(pbuf), pstride, _ptr(low), _ptr(up), bstride, ct(rtol), ct(atol), max_levels,
                ct(cache["tm"][0]), _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["xs"]), _ptr(cache["ws"]),
                C.c_long(batch), _ptr(value), _ptr(err), _ptr(level), _ptr(conv), _ptr(evals), nthreads)
        return dict(value=value, err=err, level=level, converged=conv.astype(bool), evals=evals)

    def batch_adaptive2d(self, fam, params, low, up, rtol, atol, max_levels, cache, nthreads: int = 0):
        low = np.ascontiguousarray(low, np.float64); up = np.ascontiguousarray(up, np.float64)
        batch = low.shape[0]
        pbuf = np.zeros(8) if params is None else np.ascontiguousarray(params, np.float64).reshape(-1)
        value = np.zeros(batch); err = np.zeros(batch)
        level = np.zeros(batch, np.int32); conv = np.zeros(batch, np.int32); evals = np.zeros(batch, np.int64)
        self.lib.fts_or_batch_adaptive2d_f64(
            fam, _ptr(pbuf), 0, _ptr(low), _ptr(up), 2, C.c_double(rtol), C.c_double(atol), max_levels,
            C.c_double(cache["tm"][0]), _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["xs"]), _ptr(cache["ws"]),
            C.c_long(batch), _ptr(value), _ptr(err), _ptr(level), _ptr(conv), _ptr(evals), nthreads)
        return dict(value=value, err=err, level=level, converged=conv.astype(bool), evals=evals)

    # ------------------------------------------------------------------ Double64 stand-in ---
    def dd_join(self, hi, lo):
        hi = np.ascontiguousarray(hi, np.float64).reshape(-1); lo = np.ascontiguousarray(lo, np.float64).reshape(-1)
        q = np.zeros(hi.size, "V16")
        self.lib.fts_or_dd_join(_ptr(hi), _ptr(lo), _ptr(q), C.c_long(hi.size))
        return q

    def dd_split(self, q):
        q = np.ascontiguousarray(q)
        hi = np.zeros(q.size); lo = np.zeros(q.size)
        self.lib.fts_or_dd_split(_ptr(q), _ptr(hi), _ptr(lo), C.c_long(q.size))
        return hi, lo

    def q_from_str(self, s: str):
        q = np.zeros(1, "V16")
        self.lib.fts_or_q_from_str(s.encode(), _ptr(q))
        return q

    def q_to_str(self, q) -> str:
        buf = C.create_string_buffer(80)
        self.lib.fts_or_q_to_str(_ptr(np.ascontiguousarray(q)), buf, 80)
        return buf.value.decode()

    def batch_adaptive1d_cmpl_dd(self, fam, params_q, pstride, low_q, up_q, bstride, rtol_q, atol_q,
                                 max_levels, cache, nthreads: int = 0):
        """Double64 stand-in (__float128) batch of adaptive_integrate_1D_cmpl; all arrays are V16."""
        batch = low_q.size if bstride else (params_q.size // max(pstride, 1) if pstride else 1)
        value = np.zeros(batch, "V16"); err = np.zeros(batch, "V16")
        level = np.zeros(batch, np.int32); conv = np.zeros(batch, np.int32); evals = np.zeros(batch, np.int64)
        self.lib.fts_or_batch_adaptive1d_cmpl_dd(
            fam, _ptr(params_q), pstride, _ptr(low_q), _ptr(up_q), bstride, _ptr(rtol_q), _ptr(atol_q),
            max_levels, _ptr(cache["tm"]), _ptr(cache["ix"]), _ptr(cache["iw"]), _ptr(cache["ic"]),
            _ptr(cache["xs"]), _ptr(cache["ws"]), _ptr(cache["cs"]), C.c_long(batch), _ptr(value), _ptr(err),
            _ptr(level), _ptr(conv), _ptr(evals), nthreads)
        return dict(value=value, err=err, level=level, converged=conv.astype(bool), evals=evals)


# ---------------------------------------------------------------------- `@turbo` twins ---------
class OracleAvx:
    """CPU twin of the reference's LoopVectorization kernels (oracle/fts_oracle_avx.c: SIMD
    reductions + vector libm, -O3 -march=native -ffast-math): the timed CPU baseline of bench.py."""

    def __init__(self, wide: bool = False):
        """wide=True: the build with 512-bit vectors forced (-mprefer-vector-width=512)"""
        path = os.path.join(_BUILD, f"libfts_oracle_avx{'512' if wide else ''}_{native_tag()}.so")
        if not os.path.exists(path):
            build(native=True)
        self.lib = C.CDLL(path)
        vp = C.c_void_p
        self.lib.fts_or_avx_batch_adaptive1d_f64.restype = C.c_int
        self.lib.fts_or_avx_batch_adaptive1d_f64.argtypes = [C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                                             vp, vp, vp, vp, C.c_long, vp, vp, vp, C.c_int]
        self.lib.fts_or_avx_batch_integrate1d_f64.restype = C.c_int
        self.lib.fts_or_avx_batch_integrate1d_f64.argtypes = [C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_double, C.c_long, vp, C.c_int]

    def num_threads(self) -> int:
        return self.lib.fts_or_avx_num_threads()

    def vector_bits(self) -> int:
        """widest vector ISA the build may use (gcc still prefers 256-bit unless wide=True)"""
        return self.lib.fts_or_avx_vector_bits()

    def batch_adaptive1d(self, fam, params, low, up, rtol, atol, max_levels, cache, nthreads: int = 0):
        """adaptive_integrate_1D_avx (src/adaptive.jl:94-133) looped over a batch (README.md:96-106)"""
        params = np.ascontiguousarray(params, np.float64).reshape(-1)
        low = np.ascontiguousarray(low, np.float64).reshape(-1); up = np.ascontiguousarray(up, np.float64).reshape(-1)
        batch = max(params.size, low.size)
        value = np.zeros(batch); level = np.zeros(batch, np.int32); evals = np.zeros(batch, np.int64)
        rc = self.lib.fts_or_avx_batch_adaptive1d_f64(fam, _ptr(params), 1 if params.size > 1 else 0, _ptr(low), _ptr(up), 1 if low.size > 1 else 0,
                                                      rtol, atol, max_levels, float(cache["tm"][0]), _ptr(cache["ix"]), _ptr(cache["iw"]),
                                                      _ptr(cache["xs"]), _ptr(cache["ws"]), batch, _ptr(value), _ptr(level), _ptr(evals), nthreads)
        if rc:
            raise ValueError("the @turbo twin covers the headline families only (F1_GAUSS, F1_EXP)")
        return dict(value=value, level=level, evals=evals)

    def batch_integrate1d(self, fam, params, low, up, x, w, h, nthreads: int = 0):
        """integrate1D_avx (src/integrate1D.jl:138-147) over a batch"""
        low = np.ascontiguousarray(low, np.float64).reshape(-1); up = np.ascontiguousarray(up, np.float64).reshape(-1)
        params = None if params is None else np.ascontiguousarray(params, np.float64).reshape(-1)
        batch = max(low.size, params.size if params is not None else 1)
        value = np.zeros(batch)
        x = np.ascontiguousarray(x, np.float64); w = np.ascontiguousarray(w, np.float64)
        rc = self.lib.fts_or_avx_batch_integrate1d_f64(fam, _ptr(params) if params is not None else None, 1 if (params is not None and params.size > 1) else 0,
                                                       _ptr(low), _ptr(up), 1 if low.size > 1 else 0, _ptr(x), _ptr(w), x.size, float(h), batch, _ptr(value), nthreads)
        if rc:
            raise ValueError("the @turbo twin covers the headline families only (F1_GAUSS, F1_EXP)")
        return value


# ---------------------------------------------------------------------- optimal-spacing grids --
def hopt_ref(N: int, d: float = np.pi / 2) -> float:
    """hopt(Float64, N, d) = (2/N) * lambertw(2dN)   src/core.jl:289-291 (LambertW.jl -> scipy)"""
    from scipy.special import lambertw
    return float(2.0 / N * lambertw(2.0 * d * N).real)


def tanhsinh_opt_ref(N: int, d: float = np.pi / 2, tm: float = None):
    """tanhsinh_opt(Float64, N, d)   src/core.jl:276-287: t = h:h:tm, x = ordinate.(t), w = weight.(t)
    with ordinate/weight as in src/core.jl:126-185 (clamp to prevfloat(1); weight 0 beyond 700)."""
    h = hopt_ref(N, d)
    k = int(np.floor(tm / h * (1 + 1e-12)))
    t = h * np.arange(1, k + 1)
    u = (np.pi / 2) * np.sinh(t)
    lim = np.nextafter(1.0, 0.0)
    x = np.clip(np.tanh(u), -lim, lim)
    with np.errstate(over="ignore"):
        w = np.where(np.abs(u) > 700.0, 0.0, (np.pi / 2) * np.cosh(t) / np.cosh(u) ** 2)
    return x, w, h
