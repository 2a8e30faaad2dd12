"""CPU tests of the host-side mirror of the reference API (no compute calls): table generation
against the oracle, tolerance resolution, cache memoisation and depth guard, argument checking,
quad_split fan-out geometry, NVRTC compilation (needs no GPU)."""
import ctypes as C
import math

import numpy as np
import pytest


def test_tables_match_oracle_bitwise(fts, oracle):
    for N in (8, 79, 80, 256):
        x, w, h = fts.tanhsinh(np.float64, N)
        xo, wo, ho = oracle.tanhsinh("f64", N)
        assert np.array_equal(x, xo) and np.array_equal(w, wo) and h == ho
    x, w, h = fts.tanhsinh(np.float32, 40)
    xo, wo, ho = oracle.tanhsinh("f32", 40)
    assert x.dtype == np.float32 and np.array_equal(x, xo) and np.array_equal(w, wo) and h == ho
    xd, wd, hd = fts.tanhsinh(80)  # tanhsinh(N) == tanhsinh(Float64, N)   core.jl:274
    x64, w64, h64 = fts.tanhsinh(np.float64, 80)
    assert np.array_equal(xd, x64) and hd == h64
    for T, sfx in ((np.float64, "f64"), (np.float32, "f32")):
        for D in (1, 2, 3):
            assert fts.tmax(T, D) == oracle.tmax(sfx, D)
            assert fts.t_w_max(T, D) == oracle.t_w_max(sfx, D)
        assert fts.t_x_max(T) == oracle.t_x_max(sfx)


def test_caches_match_oracle_bitwise(fts, oracle):
    c = fts.adaptive_cache_1D(np.float64, max_levels=10)
    co = oracle.cache1d("f64", 10)
    assert np.array_equal(c._xs, co["xs"]) and np.array_equal(c._ws, co["ws"]) and c.tm == co["tm"][0]
    assert np.array_equal(c.initial_x, co["ix"]) and np.array_equal(c.initial_w, co["iw"])
    cc = fts.adaptive_cache_1D(np.float64, max_levels=10, complement=True)
    cco = oracle.cache1d("f64", 10, True)
    assert np.array_equal(cc._cs, cco["cs"]) and np.array_equal(cc.initial_c, cco["ic"]) and cc.tm == cco["tm"][0]
    for D in (2, 3):
        t = (fts.adaptive_cache_2D if D == 2 else fts.adaptive_cache_3D)(np.float64, max_levels=6)
        to = oracle.cache_tensor("f64", D, 6)
        assert np.array_equal(t._xs, to["xs"]) and np.array_equal(t._ws, to["ws"]) and t.tm == to["tm"][0]
    t32 = fts.adaptive_cache_3D(np.float32, max_levels=4)
    assert np.array_equal(t32._xs, oracle.cache_tensor("f32", 3, 4)["xs"])
    dd = fts.adaptive_cache_1D(fts.Double64, max_levels=6, complement=True)
    hi, lo = oracle.dd_split(oracle.cache1d("dd", 6, True)["cs"])
    assert np.array_equal(dd._cs["hi"], hi) and np.array_equal(dd._cs["lo"], lo)
    c2 = fts.adaptive_cache_2D_cmpl(np.float64, max_levels=5)
    c2o = oracle.cache_tensor_cmpl("f64", 5)
    assert np.array_equal(c2._xs, c2o["xs"]) and np.array_equal(c2._cs, c2o["cs"]) and c2.tm == c2o["tm"][0]


def test_cache_api_semantics(fts):
    # /root/reference/test/caches.jl:1-30
    c1 = fts.adaptive_cache_1D(np.float64, max_levels=3)
    assert c1 is fts.adaptive_cache_1D(np.float64, max_levels=3)             # memoised (caches.jl:99-107)
    c1c = fts.adaptive_cache_1D(np.float64, max_levels=3, complement=True)
    assert c1 is not c1c
    c2 = fts.adaptive_cache_2D(np.float64, max_levels=3)
    c3 = fts.adaptive_cache_3D(np.float64, max_levels=2)
    assert len(c1.xs) == len(c1.ws) == len(c1.cs) == 3 and len(c2.xs) == 3 and len(c3.xs) == 2
    assert [len(v) for v in c1.xs] == [2, 4, 8] and [len(v) for v in c2.xs] == [4, 8, 16]
    assert c1.tm == fts.tmax(np.float64) and c1c.tm == fts.t_w_max(np.float64, 1)
    # depth guard raises before anything touches the device (caches.jl:25-29, test/caches.jl:24-29)
    with pytest.raises(fts.ArgumentError, match=r"supports 3 levels, but `max_levels=4` was requested"):
        fts.adaptive_integrate_1D(np.float64, fts.builtin("exp"), -1.0, 1.0, cache=c1, max_levels=4)
    with pytest.raises(fts.ArgumentError):
        fts.adaptive_integrate_2D(np.float64, fts.builtin("exp_2d"), [-1.0, -1.0], [1.0, 1.0], cache=c2, max_levels=4)
    with pytest.raises(fts.ArgumentError):
        fts.adaptive_integrate_3D(np.float64, fts.builtin("exp_3d"), [-1.0] * 3, [1.0] * 3, cache=c3, max_levels=3)
    with pytest.raises(fts.ArgumentError, match="nonnegative"):
        fts.adaptive_cache_1D(np.float64, max_levels=-1)


def test_resolve_tolerances(fts):
    # /root/reference/src/core.jl:32-44
    from fts_b200.api import _resolve_tolerances
    r, a = _resolve_tolerances(np.float64)
    assert r[0] == math.sqrt(np.finfo(np.float64).eps) and a[0] == 0.0
    r, a = _resolve_tolerances(np.float64, None, 1e-8)
    assert r[0] == 0.0 and a[0] == 1e-8
    r, a = _resolve_tolerances(np.float32, 1e-5, 0)
    assert r.dtype == np.float32 and r[0] == np.float32(1e-5)
    r, a = _resolve_tolerances(np.float32)
    assert r[0] == np.float32(math.sqrt(np.finfo(np.float32).eps))
    with pytest.raises(fts.ArgumentError, match="`rtol` must be nonnegative"):
        _resolve_tolerances(np.float64, -1e-3, 0)
    with pytest.raises(fts.ArgumentError, match="`atol` must be nonnegative"):
        _resolve_tolerances(np.float64, 1e-3, -1.0)
    r, a = _resolve_tolerances(fts.Double64, "1e-28", 0)
    assert r.dtype == fts.Double64 and abs(r["hi"][0] - 1e-28) < 1e-43 and r["lo"][0] != 0.0


def test_argument_checking_before_launch(fts):
    x, w, h = fts.tanhsinh(np.float64, 16)
    with pytest.raises(fts.DimensionMismatch, match="low must have length 2"):   # integrate2D.jl:149
        fts.integrate2D(fts.builtin("exp_2d"), [0.0], [1.0, 2.0], x, w, h)
    with pytest.raises(fts.DimensionMismatch, match="up must have length 2"):    # integrate2D.jl:150
        fts.integrate2D(fts.builtin("exp_2d"), [0.0, 0.0], [1.0], x, w, h)
    with pytest.raises(fts.DimensionMismatch):                                   # integrate3D.jl:278-279
        fts.integrate3D_avx(fts.builtin("exp_3d"), [0.0, 0.0], [1.0, 1.0, 1.0], x, w, h)
    with pytest.raises(fts.FtsError, match="Higher than 3D"):                    # quad.jl:88
        fts.quad(fts.builtin("exp_3d"), np.zeros(4), np.ones(4))
    with pytest.raises(fts.DimensionMismatch, match="same length"):              # quad.jl:70
        fts.quad(fts.builtin("exp_2d"), [0.0, 0.0], [1.0, 1.0, 1.0])
    with pytest.raises(TypeError, match="Integrand"):
        fts.quad(lambda t: t, 0.0, 1.0)
    with pytest.raises(fts.ArgumentError, match="needs params"):
        fts.quad(fts.builtin("gauss"), -1.0, 1.0)
    with pytest.raises(fts.DimensionMismatch):
        fts.quad_cmpl(fts.builtin("exp"), -1.0, 1.0)


def test_quad_split_fanout_geometry(fts, monkeypatch):
    """sub-boxes, their order and the tolerance split of quad.jl:164-281"""
    from fts_b200 import api
    seen = {}

    def fake_quad(f, low, up, **kw):
        seen.update(low=np.array(low), up=np.array(up), kw=kw)
        return np.arange(1, np.array(low).shape[0] + 1, dtype=np.float64)

    monkeypatch.setattr(api, "quad", fake_quad)
    v = api.quad_split(fts.builtin("invsqrtabs"), 0.25, -1.0, 2.0, rtol=1e-8, atol=1e-10, params=[1.0])
    assert np.array_equal(seen["low"], [-1.0, 0.25]) and np.array_equal(seen["up"], [0.25, 2.0])
    assert seen["kw"]["rtol"] == 0.5e-8 and seen["kw"]["atol"] == 0.5e-10 and seen["kw"]["max_levels"] == 16
    assert v == 3.0
    v = api.quad_split(fts.builtin("invsqrtabs_2d"), [0.0, 0.5], [-1.0, -2.0], [1.0, 2.0], rtol=1e-8)
    assert np.array_equal(seen["low"], [[-1, -2], [0, -2], [-1, 0.5], [0, 0.5]])       # q1..q4, quad.jl:212-215
    assert np.array_equal(seen["up"], [[0, 0.5], [1, 0.5], [0, 2], [1, 2]])
    assert seen["kw"]["rtol"] == 0.25e-8 and seen["kw"]["max_levels"] == 8 and v == 10.0
    api.quad_split(fts.builtin("invsqrtabs_3d"), [0.0] * 3, [-1.0] * 3, [1.0] * 3, rtol=8e-6, atol=8.0)
    assert seen["low"].shape == (8, 3) and seen["kw"]["rtol"] == 1e-6 and seen["kw"]["atol"] == 1.0 and seen["kw"]["max_levels"] == 5
    # k fastest (quad.jl:262-264): second sub-box differs from the first in z only
    assert np.array_equal(seen["low"][1], [-1, -1, 0]) and np.array_equal(seen["low"][4], [0, -1, -1])
    # default rtol stays `nothing` (quad.jl:169)
    api.quad_split(fts.builtin("invsqrtabs"), 0.0, params=[1.0])
    assert seen["kw"]["rtol"] is None and np.array_equal(seen["low"], [-1.0, 0.0])


def test_nvrtc_compiles_user_integrands_without_a_gpu(fts):
    from fts_b200 import _lib
    f = fts.cuda_integrand("return exp(-p[0]*x*x);", "1d", nparams=1)
    size, nk = C.c_size_t(), C.c_int()
    _lib.check(_lib.lib.fts_integrand_cubin_size(f._h, C.byref(size), C.byref(nk)))
    assert size.value > 10000 and nk.value >= 10 and f.kind == fts.KIND_1D and not f.is_builtin
    g = fts.cuda_integrand("return T(1)/sqrt(bmx*ymc);", "2d_cmpl")
    assert g.kind == fts.KIND_2D_CMPL
    with pytest.raises(fts.FtsError) as ei:
        fts.cuda_integrand("return not_a_function(x);", "1d")
    assert ei.value.code == 5 and "not_a_function" in str(ei.value)


def test_double64_helpers(fts):
    import mpmath
    mpmath.mp.dps = 40
    a = fts.to_dd(["3.14159265358979323846264338327950288", "1e-30"])
    back = fts.dd_to_mpf(a)
    assert abs(back[0] - mpmath.pi) < mpmath.mpf("1e-31")
    assert abs(back[1] - mpmath.mpf("1e-30")) < mpmath.mpf("1e-60")
    b = fts.to_dd([1.5, -2.0])
    assert np.array_equal(b["hi"], [1.5, -2.0]) and np.array_equal(b["lo"], [0.0, 0.0])


def test_optimal_spacing_grids(fts, oracle):
    """tanhsinh_opt / hopt / hmax (core.jl:276-302; KATs of test/core.jl:59-69) against the numpy
    restatement in oracle/fts_oracle.py"""
    import fts_oracle
    for N in (80, 81, 16, 1000):
        x, w, h = fts.tanhsinh_opt(np.float64, N)
        assert len(x) == len(w) and h > 0                                  # test/core.jl:60-66
        xr, wr, hr = fts_oracle.tanhsinh_opt_ref(N, tm=float(fts.tmax(np.float64)))
        assert abs(h - hr) <= 2e-16 * hr and len(x) == len(xr)
        assert np.max(np.abs(x - xr)) <= 4e-16 and np.max(np.abs(w - wr) / np.maximum(wr, 1e-300)) <= 5e-14
        assert fts.hopt(np.float64, N) == h and fts.hmax(np.float64, N) == fts.tmax(np.float64) / (N // 2)
    # W(z) e^{W(z)} = z in every precision
    for T, tol in ((np.float32, 3e-7), (np.float64, 5e-16)):
        hh = float(fts.hopt(T, 80))
        wv = hh * 80 / 2
        assert abs(wv * math.exp(wv) - math.pi * 80) <= tol * math.pi * 80 * 8
    hi, lo = fts.hopt(fts.Double64, 80)
    assert abs(hi - float(fts.hopt(np.float64, 80))) <= 1e-16 and abs(lo) < 1e-17
    # the grid integrates: exp on [-1,1] (spacing is optimal for N points, error ~1e-15)
    x, w, h = fts.tanhsinh_opt(np.float64, 80)
    val = h * (math.pi / 2 + float(np.sum(w * (np.exp(x) + np.exp(-x)))))
    assert abs(val - (math.e - 1 / math.e)) < 1e-14
