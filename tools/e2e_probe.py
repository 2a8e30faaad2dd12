"""Per-step distribution of the end-to-end (host buffers) time of config B — run once per value of
FTS_B200_CHUNKS (the variable is read once per process):
    for c in 1 2 4 6 8; do FTS_B200_CHUNKS=$c python tools/e2e_probe.py; done
"""
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
    h_val = torch.empty(n, dtype=torch.float64).pin_memory()

    def step():
        fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=h_alpha.data_ptr(),
                                params_stride=1, rtol=bench.RTOL, atol=bench.ATOL, max_levels=bench.MAX_LEVELS, batch=n,
                                value=h_val.data_ptr(), order="reference", options=L.OPT_QUAD_EDGE_RULES)

    out = {"chunks": os.environ.get("FTS_B200_CHUNKS", "default")}
    for with_sampler in (False, True):
        sampler = None
        if with_sampler:
            sampler = bench.ClockSampler(0)
            sampler.start()
        for _ in range(5):
            step()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            step()
            ts.append((time.perf_counter() - t0) * 1e3)
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=1.0)
        ts = np.sort(np.array(ts))
        out["sampler" if with_sampler else "plain"] = {"mean_ms": float(ts.mean()), "min_ms": float(ts[0]), "p50_ms": float(ts[100]),
                                                        "p90_ms": float(ts[180]), "max_ms": float(ts[-1])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
