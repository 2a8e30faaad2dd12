#!/usr/bin/env python
"""BASELINE config D across GPUs: ONE adaptive 3D integral sharded over all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_sharded.py [--levels 8] [--integrand log_3d] [--reps 5]

Each rank sums its share of every level's new cells, the ranks all-gather their compensated (s, c)
partials (2 doubles per rank per level, NCCL), and every rank runs the same device-side stop
test.  Rank 0 prints one JSON line: value, level, device time (max over ranks), Geval/s.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fts_b200 as fts  # noqa: E402
from fts_b200 import distributed as fd  # noqa: E402
from fts_b200.api import evals_for_level  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--integrand", default="log_3d")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rtol", type=float, default=0.0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fts.init(local)
    f = fts.builtin(args.integrand)
    cache = fts.adaptive_cache_3D(np.float64, max_levels=args.levels)
    cache.handle
    low, up = [-1.0] * 3, [1.0] * 3
    exact = {"log_3d": (2 * math.log(2) - 2) ** 3, "exp_3d": (math.e - 1 / math.e) ** 3, "invsqrt_3d": math.pi ** 3,
             "rat_3d": (2 * math.atan(5.0) / 5) * (2 * math.atan(4.0) / 4) * (2 * math.atan(3.0) / 3)}[args.integrand]
    times = []
    for it in range(args.reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fd.adaptive_integrate_sharded(np.float64, f, low, up, rtol=args.rtol, atol=0.0, max_levels=args.levels, cache=cache, warn=False, exchange=args.exchange)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 2:
            times.append(float(t.item()))
    vals = torch.tensor([float(r.value[0])], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        evals = int(evals_for_level(3, r.levels)[0])
        ms = min(times)
        print(json.dumps({"case": "D-sharded", "exchange": args.exchange, "integrand": args.integrand, "n_gpus": world, "max_levels": args.levels, "level_reached": int(r.levels[0]),
                          "value": float(r.value[0]), "rel_err_vs_exact": abs(float(r.value[0]) - exact) / abs(exact),
                          "identical_on_all_ranks": bool(all(float(v.item()) == float(r.value[0]) for v in allv)), "evals": evals, "ms": ms,
                          "ms_median": sorted(times)[len(times) // 2], "gevals_per_s": evals / ms / 1e6}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
