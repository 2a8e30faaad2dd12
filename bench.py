#!/usr/bin/env python
"""bench.py — headline benchmark of the fts_b200 hot path (BASELINE.json configs[1]).

Workload "B": 2^20 adaptive integrals of exp(-a x^2) on [-1,1] (a swept over [0.5, 50]),
`quad` semantics (reference summation order, rtol=1e-10, atol=0, max_levels=16), Float64.
One step = one pass of the hot path over the whole batch.  Per rank (one process per GPU) the
batch is fixed, so scaling over GPUs is weak and needs no collective (independent integrals).

  python bench.py [--gpus N] [--steps K] [--warmup W]      this repo's CUDA path
  python bench.py --impl reference ...                      the reference's CPU algorithm
                                                            (oracle port, all host threads)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "batched_adaptive_integrals_per_s_fp64"
UNIT = "integrals/s"
FLOP_PER_EVAL = 34.5  # 2 MUL (a*x^2) + 29 (exp) + 3.5 (map, pair add, weight FMA); FMA = 2
RTOL, ATOL, MAX_LEVELS = 1e-10, 0.0, 16
FP64_THEORETICAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2


PERMUTE = False  # --permute: the same alpha set shuffled with seed 0 (SURVEY §8(d) names both orders)


def workload(batch_log2: int):
    n = 1 << batch_log2
    alpha = 0.5 + 49.5 * np.arange(n, dtype=np.float64) / (n - 1)  # SURVEY §8(d) synthetic input B
    if PERMUTE:
        alpha = alpha[np.random.default_rng(0).permutation(n)]
    return n, alpha


def config(batch_log2: int, extra=None):
    c = {"workload": f"B: 2^{batch_log2} adaptive quad of exp(-a*x^2) on [-1,1], a in [0.5,50] ({'permuted, seed 0' if PERMUTE else 'sorted sweep'}), rtol=1e-10, atol=0, "
                     f"max_levels=16, Float64, reference summation order (quad semantics)",
         "batch_per_gpu": 1 << batch_log2, "l2_policy": "256 MiB L2 flush between timed steps (inputs are 8 MiB < L2)"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------- clocks -------
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed regions run"""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.stop_flag, self.active = index, period, [], False, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((self.active, sm, reasons, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        load = [s for s in self.samples if s[0]] or self.samples
        sm = sorted(s[1] for s in load)
        mask = 0
        for s in load:
            mask |= s[2]
        names = {getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                 getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
                 getattr(nv, "nvmlClocksThrottleReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting"}
        reasons = [n for bit, n in names.items() if mask & bit]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples_under_load": len(load),
                "power_w_max": max(s[3] for s in load)}


# ------------------------------------------------------------------------------- CPU arm ------
def cpu_reference(batch_log2: int, steps: int, warmup: int, sample_log2: int, bounded: bool):
    """The reference's own algorithm on the host cores: adaptive_integrate_1D looped over the
    batch (README.md:96-106) — oracle port compiled -O3 -march=native, OpenMP over the batch."""
    from fts_oracle import FAMILIES, Oracle, build
    try:
        build(native=True)
        orc = Oracle("native")
        variant = "-O3 -march=native"
    except Exception:
        orc = Oracle("fast")
        variant = "-O3 -march=x86-64-v3"
    n, alpha = workload(batch_log2)
    stride = 1 << max(0, batch_log2 - sample_log2)
    sample = np.ascontiguousarray(alpha[::stride])
    cache = orc.cache1d("f64", MAX_LEVELS)
    cores = orc.num_threads()
    times, evals = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = orc.batch_adaptive1d("f64", FAMILIES["F1_GAUSS"], sample, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            evals = int(out["evals"].sum())
        if bounded and sum(times) > 20.0:
            break
    per_step = sum(times) / len(times)
    # the same on ONE core (SURVEY 8d asks for both), on a 16x sparser sample
    one = np.ascontiguousarray(sample[::16]) if sample.size >= 1024 else sample
    orc.batch_adaptive1d("f64", FAMILIES["F1_GAUSS"], one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
    t0 = time.perf_counter()
    orc.batch_adaptive1d("f64", FAMILIES["F1_GAUSS"], one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
    one_core = one.size / (time.perf_counter() - t0)
    return {"value": sample.size / per_step, "unit": UNIT, "cores": cores, "kind": "port", "value_1core": one_core,
            "sample": f"every {stride}th integral of workload B ({sample.size} integrals, {evals} evals) x {len(times)} passes; "
                      f"oracle/fts_oracle.c ({variant}, OpenMP over the batch, IEEE, no reassociation)",
            "gevals_per_s": evals / per_step / 1e9, "ms_per_step": per_step * 1e3, "steps": len(times)}


# ------------------------------------------------------------------------------- main ---------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-log2", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--permute", action="store_true", help="shuffle the alpha sweep (seed 0): lanes of a warp then stop at different levels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global PERMUTE
    PERMUTE = args.permute
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        base = cpu_reference(args.batch_log2, args.steps, args.warmup, sample_log2=17, bounded=True)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
                "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config(args.batch_log2, {"cpu_sample": base["sample"]}),
                "cpu_baseline": {k: base[k] for k in ("value", "value_1core", "unit", "cores", "kind", "sample")},
                "gevals_per_s": base["gevals_per_s"],
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import fts_b200 as fts
    from fts_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fts_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fts.init(local_rank)
    dev = torch.device("cuda", local_rank)

    n, alpha_np = workload(args.batch_log2)
    cache = fts.adaptive_cache_1D(np.float64, max_levels=MAX_LEVELS)
    gauss = fts.builtin("gauss")
    alpha = torch.from_numpy(alpha_np).to(dev)
    lo = torch.tensor([-1.0], dtype=torch.float64, device=dev)
    hi = torch.tensor([1.0], dtype=torch.float64, device=dev)
    value = torch.empty(n, dtype=torch.float64, device=dev)
    err = torch.empty(n, dtype=torch.float64, device=dev)
    levels = torch.empty(n, dtype=torch.int32, device=dev)
    flags = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        fts.adaptive_batch_dev(cache, gauss, low=lo.data_ptr(), up=hi.data_ptr(), bounds_stride=0, params=alpha.data_ptr(), params_stride=1,
                               rtol=RTOL, atol=ATOL, max_levels=MAX_LEVELS, batch=n, value=value.data_ptr(), err=err.data_ptr(),
                               levels=levels.data_ptr(), flags=flags.data_ptr(), stream=stream, order="reference",
                               options=L.OPT_QUAD_EDGE_RULES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if os.environ.get("FTS_BENCH_NO_SAMPLER"):
        sampler.ok = False  # diagnostic: run without the NVML polling thread
    sampler.start()

    # measured FP64 pipe peak (register-resident DFMA chains) — the roofline denominator
    peak = fts.measure_fma_peak(np.float64, 8192)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    evals_per_step = int(fts.evals_for_level(1, levels.cpu().numpy()).sum())
    conv = int((flags & L.FLAG_CONVERGED).ne(0).sum().item())
    lv_hist = torch.bincount(levels, minlength=MAX_LEVELS + 1).cpu().tolist()

    # ---- timed region 1: inputs resident in HBM, K steps, CUDA events per step on the launch stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    fts.launch_count(reset=True)
    barrier()
    sampler.active = True
    wall0 = time.perf_counter()
    for e0, e1 in ev:
        flush.zero_()  # L2 flush, outside the event pair
        e0.record()
        step()
        e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.active = False
    launches = fts.launch_count()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    dev_ms = sum(step_ms)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    total_integrals = n * world
    value_ips = total_integrals / (ms_per_step * 1e-3)
    gevals = evals_per_step * world / (ms_per_step * 1e-3) / 1e9

    # ---- timed region 2: end to end through the host-buffer C ABI call (pinned host memory)
    h_alpha = torch.from_numpy(alpha_np).pin_memory()
    h_lo = torch.tensor([-1.0], dtype=torch.float64).pin_memory()
    h_hi = torch.tensor([1.0], dtype=torch.float64).pin_memory()
    h_val = torch.empty(n, dtype=torch.float64).pin_memory()

    def e2e_step():
        fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=h_alpha.data_ptr(),
                                params_stride=1, rtol=RTOL, atol=ATOL, max_levels=MAX_LEVELS, batch=n, value=h_val.data_ptr(),
                                order="reference", options=L.OPT_QUAD_EDGE_RULES)

    # Warm-up of the host path: >= W steps and >= 2 s.  On a box that has just idled (the reference
    # arm keeps only the CPU busy for a minute) the first seconds of PCIe traffic run at 0.4-0.6 ms per
    # step, then settle at 0.32-0.33 ms (link / host power states; measured).
    e2e_steps = max(5, min(args.steps, 200))
    t_warm, n_warm = time.perf_counter(), 0
    while n_warm < args.warmup or (time.perf_counter() - t_warm < 2.0 and n_warm < 20000):
        e2e_step()
        n_warm += 1
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    e2e_marks = [t0]
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns when the results are in h_val
        e2e_marks.append(time.perf_counter())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_each = np.diff(np.array(e2e_marks)) * 1e3
    sampler.active = False
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_val = total_integrals * e2e_steps / e2e_s
    e2e_ok = bool(np.allclose(h_val.numpy(), value.cpu().numpy(), rtol=0, atol=0))

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    achieved_tflops = evals_per_step / (ms_per_step * 1e-3) * FLOP_PER_EVAL / 1e12  # per GPU (max-over-ranks time)
    peak_tflops = peak["tflops"]
    hbm_bytes = n * (8 + 8 + 8 + 4 + 4)  # a in; value, err, level, flags out
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
    roofline = {"bound": "fp64_pipe", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                "traffic": 8406528,  # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/r1_adaptive1d_tpi_*_ncu_full.md): the 8 MiB of alpha; outputs stay in L2

                "peak_source": "in-run register-resident DFMA chain microbenchmark (fts_measure_fma_peak); MEASURED_PEAKS.json has no FP64 figure",
                "peak_theoretical": FP64_THEORETICAL_TFLOPS, "frac_of_theoretical": achieved_tflops / FP64_THEORETICAL_TFLOPS,
                "flop_per_eval": FLOP_PER_EVAL, "evals_per_launch": evals_per_step, "kernel": "k_adaptive1d_tpi<double,F1Gauss,false>",
                "kernel_ms": ms_per_step,
                "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / (ms_per_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "note": "not the binding roofline: arithmetic intensity ~290 flop/B"}}
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_reference(args.batch_log2, steps=5, warmup=1, sample_log2=args.batch_log2, bounded=True)
            cpu = {k: cpu[k] for k in ("value", "value_1core", "unit", "cores", "kind", "sample", "gevals_per_s")}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
    line = {"metric": METRIC, "value": value_ips, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.batch_log2, {"parallelism": f"batch sharded over {world} GPU(s), no collective"}),
            "gevals_per_s": gevals, "evals_per_integral_mean": evals_per_step / n, "level_histogram": lv_hist, "converged": conv,
            "wall_s_timed_region": wall, "step_ms_min": min(step_ms), "step_ms_median": sorted(step_ms)[len(step_ms) // 2],
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n * 8 + 16, "d2h_bytes_per_step": n * 8, "steps": e2e_steps,
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "step_ms_min": float(e2e_each.min()), "step_ms_median": float(np.median(e2e_each)),
                    "warmup_steps": n_warm, "matches_device_path_bitwise": e2e_ok,
                    "api": "fts_adaptive_batch on pinned host buffers: the kernel reads alpha from and writes the values to host memory in place over PCIe (every step, inside the timed region); FTS_B200_ZEROCOPY=0 selects the staged 4-chunk copy pipeline"},
            "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
