#!/usr/bin/env python
"""Supplementary measurements for the BASELINE.json configs other than the headline one
(bench.py measures configs[1]).  Every number is end to end through the host-buffer C ABI
(PCIe copies included), best of `reps` after warm-up, one JSON line per case.

    python tools/bench_configs.py [--cases A,C,D,E] [--reps 10]
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"))

import numpy as np  # noqa: E402

import fts_b200 as fts  # noqa: E402
from fts_b200.api import evals_for_level  # noqa: E402

FP64_PEAK = None
FP32_PEAK = None


def best_of(fn, reps):
    fn()
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


RECORDS = []  # bench.py calls the cases in-process and reads the records here instead of stdout
QUIET = False


def emit(**kw):
    RECORDS.append(kw)
    if not QUIET:
        print(json.dumps(kw), flush=True)


def case_a(reps):
    """config A: integrate1D[_avx] of exp on [0, 1+k 2^-20], precomputed (x,w,h); flop/eval 32.5"""
    n = 1 << 20
    up = 1.0 + np.arange(n) * 2.0 ** -20
    low = np.zeros(n)
    f = fts.builtin("exp")
    for N in (80, 256, 1024):
        x, w, h = fts.tanhsinh(np.float64, N)
        for order in ("reference", "fast"):
            t, v = best_of(lambda: fts.integrate1D(f, low, up, x, w, h, order=order), reps)
            evals = n * (2 * (N // 2) + 1)
            err = float(np.max(np.abs(v - (np.exp(up) - 1)) / (np.exp(up) - 1)))
            emit(case="A", api="integrate1D" + ("_avx" if order == "fast" else ""), N=N, batch=n, ms=t * 1e3, integrals_per_s=n / t,
                 gevals_per_s=evals / t / 1e9, fp64_frac_nominal=evals / t * 32.5 / 1e12 / FP64_PEAK, max_rel_err_vs_exact=err)


def case_apin(reps):
    """config A through fts_integrate_batch on PINNED host buffers (bounds read and values written in place over PCIe)"""
    import ctypes as C
    from fts_b200 import _lib as L, api
    n = 1 << 20
    low, up, out = fts.pinned_empty(n), fts.pinned_empty(n), fts.pinned_empty(n)
    up[...] = 1.0 + np.arange(n) * 2.0 ** -20
    low[...] = 0.0
    f = fts.builtin("exp")
    for N in (80, 256, 1024):
        x, w, h = fts.tanhsinh(np.float64, N)
        handle = api._tables.get(x, w, h)
        for order in ("reference", "fast"):
            def call():
                L.check(L.lib.fts_integrate_batch(handle, f._h, api._order(order), C.c_void_p(low.ctypes.data), C.c_void_p(up.ctypes.data), 1, None, 0,
                                                  n, C.c_void_p(out.ctypes.data)))
                return out
            t, v = best_of(call, reps)
            evals = n * (2 * (N // 2) + 1)
            err = float(np.max(np.abs(v - (np.exp(up) - 1)) / (np.exp(up) - 1)))
            emit(case="A-pinned-host", api="integrate1D" + ("_avx" if order == "fast" else ""), N=N, batch=n, ms=t * 1e3, integrals_per_s=n / t,
                 gevals_per_s=evals / t / 1e9, fp64_frac_nominal=evals / t * 32.5 / 1e12 / FP64_PEAK, max_rel_err_vs_exact=err,
                 pcie_bytes=24 * n)


def case_adev(reps):
    """config A with inputs resident in HBM: kernel-only time (CUDA events), reference and FMA order"""
    import torch
    from fts_b200 import api
    dev = torch.device("cuda", 0)
    n = 1 << 20
    up_np = 1.0 + np.arange(n) * 2.0 ** -20
    up = torch.from_numpy(up_np).to(dev)
    low = torch.zeros(n, dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    f = fts.builtin("exp")
    stream = torch.cuda.current_stream().cuda_stream
    for N in (80, 256, 1024):
        x, w, h = fts.tanhsinh(np.float64, N)
        for order in ("reference", "fast"):
            ts = []
            for it in range(reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                api.integrate_batch_dev(x, w, h, f, low=low.data_ptr(), up=up.data_ptr(), bounds_stride=1, params=0, params_stride=0, batch=n,
                                        value=out.data_ptr(), stream=stream, order=order)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = min(ts)
            evals = n * (2 * (N // 2) + 1)
            err = float(np.max(np.abs(out.cpu().numpy() - (np.exp(up_np) - 1)) / (np.exp(up_np) - 1)))
            emit(case="A-device", api="integrate1D" + ("_avx" if order == "fast" else ""), N=N, batch=n, ms=ms, integrals_per_s=n / ms * 1e3,
                 gevals_per_s=evals / ms / 1e6, fp64_frac_nominal=evals / ms * 1e3 * 32.5 / 1e12 / FP64_PEAK, max_rel_err_vs_exact=err)


def case_c(reps):
    """config C: 4096 boxes, 1/sqrt((b-x)(y-c)), tensor-complement adaptive 2D; flop/eval 14.5"""
    rng = np.random.default_rng(0)
    n = 4096
    low = np.stack([-3 + 2 * rng.random(n), -3 + 2 * rng.random(n)], axis=1)
    up = np.stack([1 + 2 * rng.random(n), 1 + 2 * rng.random(n)], axis=1)
    exact = 4 * np.sqrt((up[:, 0] - low[:, 0]) * (up[:, 1] - low[:, 1]))
    f = fts.builtin("c2_invsqrt_bmx_ymc")
    for T, rtol, ml in ((np.float64, 1e-10, 8), (np.float32, 1e-5, 6)):
        cache = fts.adaptive_cache_2D_cmpl(T, max_levels=ml)
        lo, hi = low.astype(T), up.astype(T)
        t, r = best_of(lambda: fts.quad_cmpl(f, lo, hi, rtol=rtol, max_levels=ml, cache=cache, order="fast", full_output=True), reps)
        evals = int(evals_for_level(2, r.levels).sum())
        emit(case="C", dtype=np.dtype(T).name, batch=n, rtol=rtol, max_levels=ml, ms=t * 1e3, integrals_per_s=n / t, gevals_per_s=evals / t / 1e9,
             levels=np.bincount(r.levels).tolist(), converged=int(r.converged.sum()),
             max_rel_err_vs_exact=float(np.max(np.abs(r.value - exact) / exact)),
             fp64_frac_nominal=(evals / t * 14.5 / 1e12 / FP64_PEAK) if T == np.float64 else None)


def case_cdev(reps):
    """config C with inputs resident in HBM: kernel-only time (CUDA events) of the cooperative 2D path"""
    import torch
    from fts_b200 import api
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    n = int(os.environ.get("FTS_BENCH_C_BATCH", 4096))  # BASELINE: 4096 (one wave of warps); larger batches show the kernel's throughput
    low = np.stack([-3 + 2 * rng.random(n), -3 + 2 * rng.random(n)], axis=1)
    up = np.stack([1 + 2 * rng.random(n), 1 + 2 * rng.random(n)], axis=1)
    exact = 4 * np.sqrt((up[:, 0] - low[:, 0]) * (up[:, 1] - low[:, 1]))
    f = fts.builtin("c2_invsqrt_bmx_ymc")
    stream = torch.cuda.current_stream().cuda_stream
    for T, tt, rtol, ml in ((np.float64, torch.float64, 1e-10, 8), (np.float32, torch.float32, 1e-5, 6)):
        ml = int(os.environ.get("FTS_BENCH_C_MAX_LEVELS", ml))  # 5: nothing can be queued, the queue pass is not launched
        cache = fts.adaptive_cache_2D_cmpl(T, max_levels=ml)
        lo, hi = torch.from_numpy(low.astype(T)).to(dev), torch.from_numpy(up.astype(T)).to(dev)
        val = torch.empty(n, dtype=tt, device=dev)
        lev = torch.empty(n, dtype=torch.int32, device=dev)
        # K calls back to back per timed region, like the steps of bench.py: with ONE call between the events the
        # 10 us the host spends between recording the first event and reaching cudaLaunchKernel (Python, ctypes,
        # argument checks) are counted as device time of a 28 us kernel
        K = int(os.environ.get("FTS_BENCH_C_CALLS", 10))
        ts = []
        for it in range(reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                api.adaptive_batch_dev(cache, f, low=lo.data_ptr(), up=hi.data_ptr(), bounds_stride=2, params=0, params_stride=0, rtol=rtol, atol=0.0,
                                       max_levels=ml, batch=n, value=val.data_ptr(), levels=lev.data_ptr(), stream=stream, order="fast",
                                       options=1)  # FTS_OPT_QUAD_EDGE_RULES
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) / K)
        t_loop = min(ts) * 1e-3
        # how long the HOST needs to queue one call (Python, ctypes, argument checks, two launches): when this is not
        # below the device time, the loop above measures the host
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(20 * K):
            api.adaptive_batch_dev(cache, f, low=lo.data_ptr(), up=hi.data_ptr(), bounds_stride=2, params=0, params_stride=0, rtol=rtol, atol=0.0,
                                   max_levels=ml, batch=n, value=val.data_ptr(), levels=lev.data_ptr(), stream=stream, order="fast", options=1)
        host_us = (time.perf_counter() - w0) / (20 * K) * 1e6
        torch.cuda.synchronize()
        # the same K calls captured into a CUDA graph (the entry point only queues work on the caller's stream) and
        # replayed: device time per call without the host in the loop
        t_graph = None
        try:
            side = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                cs = torch.cuda.current_stream().cuda_stream
                for _ in range(K):
                    api.adaptive_batch_dev(cache, f, low=lo.data_ptr(), up=hi.data_ptr(), bounds_stride=2, params=0, params_stride=0, rtol=rtol,
                                           atol=0.0, max_levels=ml, batch=n, value=val.data_ptr(), levels=lev.data_ptr(), stream=cs, order="fast",
                                           options=1)
            tg = []
            for it in range(reps + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    tg.append(e0.elapsed_time(e1) / K)
            t_graph = min(tg) * 1e-3
        except Exception as ex:  # noqa: BLE001 — the figure is optional, the reason is reported
            print(f"[bench_configs] graph capture of config C failed: {ex}", file=sys.stderr)
        t = t_graph if t_graph is not None else t_loop
        levels = lev.cpu().numpy()
        evals = int(evals_for_level(2, levels).sum())
        emit(case="C-device", dtype=np.dtype(T).name, batch=n, ms=t * 1e3, ms_python_loop=t_loop * 1e3, ms_graph_replay=None if t_graph is None else t_graph * 1e3,
             host_us_per_call=host_us, ms_median=float(np.median(ts)), integrals_per_s=n / t,
             gevals_per_s=evals / t / 1e9, levels=np.bincount(levels).tolist(),
             max_rel_err_vs_exact=float(np.max(np.abs(val.cpu().numpy() - exact) / exact)),
             fp64_frac_nominal=(evals / t * 14.5 / 1e12 / FP64_PEAK) if T == np.float64 else None,
             # Float32: the same 14.5 flop/eval against the measured FFMA peak, and one MUFU.RSQ per evaluation against
             # the SFU rate (16 per clock and SM: 148 x 16 x 1.965 GHz = 4.65 T/s)
             fp32_frac_nominal=(evals / t * 14.5 / 1e12 / FP32_PEAK) if (T == np.float32 and FP32_PEAK) else None,
             mufu_frac=(evals / t / 4.65e12) if T == np.float32 else None,
             warp2d_levels=os.environ.get("FTS_B200_WARP2D_LEVELS", "default(5)"), calls_per_timed_region=K)


def case_d(reps):
    """config D on ONE GPU: a single adaptive 3D integral of exp(x+y+z), forced depth; flop/eval 32.1"""
    f = fts.builtin("exp_3d")
    exact = (math.e - 1 / math.e) ** 3
    for L in (5, 6, 7, 8):
        cache = fts.adaptive_cache_3D(np.float64, max_levels=L)
        t, r = best_of(lambda: fts.adaptive_integrate_3D_avx(np.float64, f, [-1.0] * 3, [1.0] * 3, rtol=0.0, atol=0.0, max_levels=L, warn=False,
                                                             cache=cache, full_output=True), max(3, reps // 2))
        evals = int(evals_for_level(3, r.levels).sum())
        emit(case="D", integrand="exp(x+y+z)", max_levels=L, level_reached=int(r.levels[0]), evals=evals, ms=t * 1e3, gevals_per_s=evals / t / 1e9,
             rel_err_vs_exact=float(abs(r.value[0] - exact) / exact), fp64_frac_nominal=evals / t * 32.1 / 1e12 / FP64_PEAK)
    # a genuinely large single integral: log(1-x)log(1-y)log(1-z) (test/families_3d.jl:64) keeps refining to max_levels
    g = fts.builtin("log_3d")
    exact_log = (2 * math.log(2) - 2) ** 3
    for L in (7, 8):
        cache = fts.adaptive_cache_3D(np.float64, max_levels=L)
        t, r = best_of(lambda: fts.adaptive_integrate_3D_avx(np.float64, g, [-1.0] * 3, [1.0] * 3, rtol=0.0, atol=0.0, max_levels=L, warn=False,
                                                             cache=cache, full_output=True), 3)
        evals = int(evals_for_level(3, r.levels).sum())
        emit(case="D", integrand="log(1-x)log(1-y)log(1-z)", max_levels=L, level_reached=int(r.levels[0]), evals=evals, ms=t * 1e3,
             gevals_per_s=evals / t / 1e9, rel_err_vs_exact=float(abs(r.value[0] - exact_log) / abs(exact_log)))


def case_dk(reps):
    """config D, kernel only: one refinement level of ONE 3D integral (k_level_grid), CUDA events"""
    import torch
    from fts_b200 import api
    dev = torch.device("cuda", 0)
    lo = torch.tensor([-1.0] * 3, dtype=torch.float64, device=dev)
    up = torch.tensor([1.0] * 3, dtype=torch.float64, device=dev)
    partial = torch.zeros(2, dtype=torch.float64, device=dev)
    cache = fts.adaptive_cache_3D(np.float64, max_levels=8)
    stream = torch.cuda.current_stream().cuda_stream
    only = os.environ.get("FTS_BENCH_DK")  # e.g. "log_3d:8": one family / level (for an ncu capture)
    fams = (("exp_3d", 32.1), ("x2y2z2", 6.0), ("log_3d", 35.0), ("invsqrt_3d", 14.5), ("rat_3d", 12.0))  # nominal flop/eval (exp 29, log 35, rsqrt 12)
    for name, flop in fams:
        if only and name != only.split(":")[0]:
            continue
        f = fts.builtin(name)
        for level in (6, 7, 8):
            if only and level != int(only.split(":")[1]):
                continue
            n = 2 << level
            evals = 7 * n ** 3 + 9 * n ** 2 + 3 * n
            ts = []
            for it in range(reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                api.sharded_level_dev(cache, f, low=lo.data_ptr(), up=up.data_ptr(), params=0, level=level, rank=0, nranks=1,
                                      partial=partial.data_ptr(), state=None, stream=stream)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = min(ts)
            emit(case="D-kernel", integrand=name, level=level, new_evals=evals, ms=ms, gevals_per_s=evals / ms / 1e6,
                 fp64_frac_nominal=(evals / ms / 1e-3 * flop / 1e12 / FP64_PEAK) if flop else None)


def case_e(reps):
    """config E: Double64 quad_cmpl, 65536 integrals, rtol 1e-28"""
    import mpmath
    mpmath.mp.dps = 40
    rng = np.random.default_rng(0)
    n = 65536
    a = -3 * rng.random(n); b = 0.5 + 3.5 * rng.random(n); c = 0.5 + 1.5 * rng.random(n)
    cache = fts.adaptive_cache_1D(fts.Double64, max_levels=16, complement=True)
    A, B, Cc = fts.to_dd(a), fts.to_dd(b), fts.to_dd(c)
    for name, params in (("c_invsqrt", Cc), ("c_log_bmx", None)):
        f = fts.builtin(name)
        t, r = best_of(lambda: fts.quad_cmpl(f, A, B, rtol="1e-28", max_levels=16, cache=cache, params=params, full_output=True), max(3, reps // 3))
        evals = int(evals_for_level(1, r.levels).sum())
        idx = [0, n // 2, n - 1]
        got = fts.dd_to_mpf(r.value[idx])
        if name == "c_invsqrt":
            want = [mpmath.mpf(float(c[i])) * mpmath.pi for i in idx]
        else:
            want = [(mpmath.mpf(float(b[i])) - mpmath.mpf(float(a[i]))) * (mpmath.log(mpmath.mpf(float(b[i])) - mpmath.mpf(float(a[i]))) - 1) for i in idx]
        err = max(float(abs(g - w_) / abs(w_)) for g, w_ in zip(got, want))
        emit(case="E", family=name, batch=n, ms=t * 1e3, integrals_per_s=n / t, gevals_per_s=evals / t / 1e9, levels=np.bincount(r.levels).tolist(),
             converged=int(r.converged.sum()), max_rel_err_vs_exact_sampled=err)


def case_edev(reps):
    """config E with inputs resident in HBM: kernel-only time (CUDA events) of the Double64 complement kernel"""
    import torch
    from fts_b200 import api
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    n = 65536
    a = -3 * rng.random(n); b = 0.5 + 3.5 * rng.random(n); c = 0.5 + 1.5 * rng.random(n)
    cache = fts.adaptive_cache_1D(fts.Double64, max_levels=16, complement=True)

    def dd_dev(v):  # [n] float64 -> [n, 2] (hi, lo) = the fts_dd layout
        return torch.from_numpy(np.stack([v, np.zeros_like(v)], axis=1)).to(dev)

    A, B, Cc = dd_dev(a), dd_dev(b), dd_dev(c)
    val = torch.empty(n, 2, dtype=torch.float64, device=dev)
    lev = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for name, params in (("c_invsqrt", Cc), ("c_log_bmx", None)):
        f = fts.builtin(name)
        ts = []
        for it in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            api.adaptive_batch_dev(cache, f, low=A.data_ptr(), up=B.data_ptr(), bounds_stride=1, params=params.data_ptr() if params is not None else 0,
                                   params_stride=1 if params is not None else 0, rtol="1e-28", atol=0, max_levels=16, batch=n, value=val.data_ptr(),
                                   levels=lev.data_ptr(), stream=stream, order="reference", options=1)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        t = min(ts) * 1e-3
        levels = lev.cpu().numpy()
        evals = int(evals_for_level(1, levels).sum())
        emit(case="E-device", family=name, batch=n, ms=t * 1e3, integrals_per_s=n / t, gevals_per_s=evals / t / 1e9, levels=np.bincount(levels).tolist())


def case_nd(reps):
    """integrateND (n-D extension): exp(sum x) on [-1,1]^D through the host-buffer call (wall clock incl. the
    synchronisation; 16 B of bounds per axis in, one value out); flop/eval = 29 (exp) + 2 D (coordinates, weight) + 7"""
    body = "T s = x[0]; for (int d = 1; d < ND; ++d) s = s + x[d]; return exp(s);"
    for D, N in ((4, 160), (5, 60), (6, 28)):
        x, w, h = fts.tanhsinh(np.float64, N, D=D)
        f = fts.cuda_integrand(body, "nd", dim=D)
        t, v = best_of(lambda: fts.integrateND(f, [-1.0] * D, [1.0] * D, x, w, h), max(3, reps // 2))
        pts = (2 * len(x) + 1) ** D
        s1 = h * (math.pi / 2 + float(np.sum(w * (np.exp(x) + np.exp(-x)))))
        emit(case="ND", D=D, N=N, points=pts, ms=t * 1e3, gevals_per_s=pts / t / 1e9, rel_err_vs_separable_sum=abs(v - s1 ** D) / s1 ** D,
             fp64_frac_nominal=pts / t * (36 + 2 * D) / 1e12 / FP64_PEAK)


def main():
    global FP64_PEAK, FP32_PEAK
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="A,C,D,E")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    fts.init(0)
    FP64_PEAK = fts.measure_fma_peak(np.float64, 8192)["tflops"]
    FP32_PEAK = fts.measure_fma_peak(np.float32, 8192)["tflops"]
    emit(case="peak", fp64_tflops_measured=FP64_PEAK, fp32_tflops_measured=FP32_PEAK)
    for c in args.cases.split(","):
        {"A": case_a, "Apin": case_apin, "Adev": case_adev, "C": case_c, "Cdev": case_cdev, "D": case_d, "Dk": case_dk, "E": case_e, "Edev": case_edev, "ND": case_nd}[c](args.reps)


if __name__ == "__main__":
    main()
