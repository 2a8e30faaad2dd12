"""Where the ~50 us between the kernel time of config B (0.268 ms) and one end-to-end call (0.319 ms) go: wall clock per
call of (a) fts_adaptive_batch_dev on device buffers + stream synchronise, (b) fts_adaptive_batch on pinned host buffers
(value + flags back), (c) the same with per-integral bounds arrays instead of two broadcast scalars (no staging copies)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fasttanhsinhquadrature.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fts_b200 as fts  # noqa: E402
from fts_b200 import _lib as L, api  # noqa: E402
import bench  # noqa: E402


def timeit(step, name, extra=None):
    t_end = time.perf_counter() + 1.5
    while time.perf_counter() < t_end:
        step()
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(json.dumps({"variant": name, "ms_median": float(np.median(ts)), "ms_min": float(ts.min()), "ms_p90": float(np.percentile(ts, 90)), **(extra or {})}), flush=True)


def main():
    fts.init(0)
    n, alpha_np = bench.workload(20)
    cache = fts.adaptive_cache_1D(np.float64, max_levels=bench.MAX_LEVELS)
    gauss = fts.builtin("gauss")
    stream = torch.cuda.current_stream().cuda_stream
    d_alpha = torch.from_numpy(alpha_np).cuda()
    d_lo, d_hi = torch.tensor([-1.0], dtype=torch.float64).cuda(), torch.tensor([1.0], dtype=torch.float64).cuda()
    d_val, d_flg = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")

    def dev_step():
        api.adaptive_batch_dev(cache, gauss, low=d_lo.data_ptr(), up=d_hi.data_ptr(), bounds_stride=0, params=d_alpha.data_ptr(), params_stride=1,
                               rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n, value=d_val.data_ptr(), flags=d_flg.data_ptr(),
                               stream=stream, order="reference", options=L.OPT_QUAD_EDGE_RULES)
        torch.cuda.current_stream().synchronize()

    timeit(dev_step, "_dev entry point on device buffers + stream synchronise")
    h_alpha = torch.from_numpy(alpha_np).pin_memory()
    hv, hf = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
    # the same entry point with pinned host pointers (unified addressing: usable on the device as they are) for some operands
    for name, pa, pv, pf in (("alpha in host memory, results on the device", h_alpha, d_val, d_flg), ("alpha on the device, value + flags in host memory", d_alpha, hv, hf),
                             ("alpha on the device, value in host memory, flags on the device", d_alpha, hv, d_flg), ("all three in host memory", h_alpha, hv, hf)):
        def mixed_step(pa=pa, pv=pv, pf=pf):
            api.adaptive_batch_dev(cache, gauss, low=d_lo.data_ptr(), up=d_hi.data_ptr(), bounds_stride=0, params=pa.data_ptr(), params_stride=1,
                                   rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n, value=pv.data_ptr(), flags=pf.data_ptr(),
                                   stream=stream, order="reference", options=L.OPT_QUAD_EDGE_RULES)
            torch.cuda.current_stream().synchronize()
        timeit(mixed_step, "_dev entry point: " + name)
    # would a LOADER that streams alpha into device memory ahead of the compute (latency-insensitive reads of the same 8 MiB)
    # help?  Emulation: alpha on the device, results in host memory, and the copy engine reads the 8 MiB of alpha from host
    # memory at the same time on a second stream
    side = torch.cuda.Stream()
    d_alpha2 = torch.empty_like(d_alpha)

    def emu_step():
        with torch.cuda.stream(side):
            d_alpha2.copy_(h_alpha, non_blocking=True)
        api.adaptive_batch_dev(cache, gauss, low=d_lo.data_ptr(), up=d_hi.data_ptr(), bounds_stride=0, params=d_alpha.data_ptr(), params_stride=1,
                               rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n, value=hv.data_ptr(), flags=hf.data_ptr(),
                               stream=stream, order="reference", options=L.OPT_QUAD_EDGE_RULES)
        torch.cuda.current_stream().synchronize()
        side.synchronize()
    timeit(emu_step, "_dev entry point: alpha on the device, results in host memory, concurrent copy-engine H2D of alpha")
    h_lo, h_hi = torch.tensor([-1.0], dtype=torch.float64).pin_memory(), torch.tensor([1.0], dtype=torch.float64).pin_memory()
    h_val, h_flg = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()

    def host_step():
        fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=h_alpha.data_ptr(), params_stride=1,
                                rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n, value=h_val.data_ptr(), flags=h_flg.data_ptr(),
                                order="reference", options=L.OPT_QUAD_EDGE_RULES)

    timeit(host_step, "host entry point, pinned buffers in place, broadcast bounds (two 8-byte staging copies)")
    h_los, h_his = torch.full((n,), -1.0, dtype=torch.float64).pin_memory(), torch.full((n,), 1.0, dtype=torch.float64).pin_memory()

    def host_step2():
        fts.adaptive_batch_host(cache, gauss, low=h_los.data_ptr(), up=h_his.data_ptr(), bounds_stride=1, params=h_alpha.data_ptr(), params_stride=1,
                                rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n, value=h_val.data_ptr(), flags=h_flg.data_ptr(),
                                order="reference", options=L.OPT_QUAD_EDGE_RULES)

    timeit(host_step2, "host entry point, per-integral bounds in place (no staging copy, 16 MiB more over PCIe)")


if __name__ == "__main__":
    main()
