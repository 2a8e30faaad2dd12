"""What bounds the end-to-end (host buffers, in place over PCIe) time of config B: the same call with different result
sets — value only (8 B per integral), value + flags (12 B, what bench.py's e2e returns), value + flags + err + levels
(24 B).  One JSON line per variant.  (One B200, PCIe 5 x16: value 0.305-0.310 ms, value + flags 0.318 ms, all four
0.66-0.77 ms: 24 MiB back are 0.43 ms of link time — the path is bound by the device-to-host direction.)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fasttanhsinhquadrature.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fts_b200 as fts  # noqa: E402
from fts_b200 import _lib as L  # noqa: E402
import bench  # noqa: E402


def main():
    fts.init(0)
    n, alpha_np = bench.workload(20)
    cache = fts.adaptive_cache_1D(np.float64, max_levels=bench.MAX_LEVELS)
    gauss = fts.builtin("gauss")
    h_alpha = torch.from_numpy(alpha_np).pin_memory()
    h_lo = torch.tensor([-1.0], dtype=torch.float64).pin_memory()
    h_hi = torch.tensor([1.0], dtype=torch.float64).pin_memory()
    h_val, h_err = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    h_flg, h_lev = torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
    variants = {
        "value": dict(params=h_alpha, value=h_val),
        "value+flags": dict(params=h_alpha, value=h_val, flags=h_flg),
        "value+flags+err+levels": dict(params=h_alpha, value=h_val, flags=h_flg, err=h_err, levels=h_lev),
    }
    for name, v in variants.items():
        kw = {k: t.data_ptr() for k, t in v.items() if k != "params"}

        def step():
            fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=v["params"].data_ptr(),
                                    params_stride=1, rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n,
                                    order="reference", options=L.OPT_QUAD_EDGE_RULES, **kw)

        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            step()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(json.dumps({"variant": name, "ms_median": float(np.median(ts)), "ms_min": float(ts.min()), "ms_p90": float(np.percentile(ts, 90))}), flush=True)


if __name__ == "__main__":
    main()
