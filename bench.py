#!/usr/bin/env python
"""bench.py — headline benchmark of the fts_b200 hot path (BASELINE.json configs[1]).

Workload "B": 2^20 adaptive integrals of exp(-a x^2) on [-1,1] (a swept over [0.5, 50]),
`quad` semantics (reference summation order, rtol=1e-10, atol=0, max_levels=16), Float64.
One step = one pass of the hot path over the whole batch.  Per rank (one process per GPU) the
batch is fixed, so scaling over GPUs is weak and needs no collective (independent integrals).

  python bench.py [--gpus N] [--steps K] [--warmup W]      this repo's CUDA path
  python bench.py --impl reference ...                      the reference's CPU algorithm
                                                            (C port of its scalar + @turbo kernels,
                                                            all host threads)
Prints ONE JSON line (rank 0).  Besides the contract's keys the line carries
  roofline      FP64-pipe roofline of the dominant kernel: model fraction, EXECUTED-instruction
                fraction and DRAM traffic (the last two from the committed ncu capture of the same
                kernel, profiles/kernel_profiles.json; stale captures are refused)
  e2e           the same metric through the host-buffer C ABI call (pinned buffers, copies inside
                the timed region): steady state, cold (driver warm-up only), and the box's
                concurrent host-link ceiling measured with plain pinned cudaMemcpyAsync
  permuted      the same workload with the parameter set shuffled (seed 0)
  sharded3d     BASELINE configs[3]: ONE large 3D integral sharded over the N GPUs (the only path
                with an exchange step), fused peer-memory exchange and NCCL all-gather variants
  other_configs BASELINE configs[2] (4096 2D complement boxes, FP64 / FP32) and configs[4] (65 536 Double64
                quad_cmpl integrals): device-resident, CUDA events (tools/bench_configs.py cases Cdev, Edev)
  cpu_baseline  the CPU port on the box's host cores (scalar = what `quad` runs; value_avx = the
                @turbo twin)
"""
import argparse
import hashlib
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
FLOP_PER_EVAL = 34.5  # SURVEY §8(d) cost MODEL: 2 MUL (a*x^2) + 29 (libm-class exp) + 3.5 (map, pair add, weight FMA); FMA = 2
RTOL, ATOL, MAX_LEVELS = 1e-10, 0.0, 16
FP64_THEORETICAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2
HEADLINE_KERNEL = "k_adaptive1d_tpi<double, F1Gauss, 0>"  # as ncu prints it
PROFILES = os.path.join(ROOT, "profiles", "kernel_profiles.json")
HEADLINE_SOURCES = ("fts_dd.cuh", "fts_math.cuh", "fts_exp_table.inc", "fts_families.cuh", "fts_kernels.cuh")  # what the headline kernel is built from

PERMUTE = False  # --permute: the same alpha set shuffled with seed 0 (SURVEY §8(d) names both orders)


def kernel_source_sha(files=HEADLINE_SOURCES) -> str:
    """hash of the device sources a profiled kernel is built from (a profile carries the file list and the hash it was taken at)"""
    h = hashlib.sha256()
    for name in files:
        path = os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200", "csrc", name)
        h.update(name.encode())
        if os.path.exists(path):
            h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


def committed_profile(kernel: str):
    """(record, error) of `kernel` in profiles/kernel_profiles.json; a capture taken at other kernel sources is refused"""
    if not os.path.exists(PROFILES):
        return None, "profiles/kernel_profiles.json is missing"
    rec = json.load(open(PROFILES)).get(kernel)
    if rec is None:
        return None, f"no committed ncu capture of {kernel}"
    now = kernel_source_sha(tuple(rec.get("source_files", HEADLINE_SOURCES)))
    if rec.get("source_sha16") != now:
        return None, (f"STALE ncu capture of {kernel}: taken at kernel sources {rec.get('source_sha16')}, current {now} "
                      f"— re-profile (tools/ncu_profile_json.py)")
    return rec, None


def workload(batch_log2: int, permute=None):
    n = 1 << batch_log2
    alpha = 0.5 + 49.5 * np.arange(n, dtype=np.float64) / (n - 1)  # SURVEY §8(d) synthetic input B
    if PERMUTE if permute is None else permute:
        alpha = alpha[np.random.default_rng(0).permutation(n)]
    return n, alpha


def config(batch_log2: int, world: int):
    """identical in both arms (the driver compares them)"""
    return {"workload": f"B: 2^{batch_log2} adaptive quad of exp(-a*x^2) on [-1,1], a in [0.5,50] ({'permuted, seed 0' if PERMUTE else 'sorted sweep'}), rtol=1e-10, atol=0, "
                        f"max_levels=16, Float64, reference summation order (quad semantics)",
            "batch_per_gpu": 1 << batch_log2, "l2_policy": "256 MiB L2 flush between timed steps (inputs are 8 MiB < L2)",
            "parallelism": f"batch sharded over {world} GPU(s), no collective"}


def host_cores() -> int:
    """host threads this process may use: the affinity mask, NOT OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
def cpu_reference(batch_log2: int, steps: int, warmup: int, budget_s: float = 20.0):
    """The reference's own algorithm on the host cores over the FULL workload, one pass per step:
    adaptive_integrate_1D looped over the batch (README.md:96-106) — (i) the scalar kernel `quad`
    runs (adaptive.jl:28-75; C port, -O3 -march=native, IEEE, no reassociation) and (ii) the
    `@turbo` twin (adaptive.jl:94-133; SIMD reduction + vector libm exp, -ffast-math), both under
    OpenMP over the batch on every core this process may use."""
    from fts_oracle import FAMILIES, Oracle, OracleAvx, build
    cores = host_cores()
    try:
        build(native=True)
        orc = Oracle("native")
        variant = "-O3 -march=native"
    except Exception:
        orc = Oracle("fast")
        variant = "-O3 -march=x86-64-v3"
    n, alpha = workload(batch_log2)
    cache = orc.cache1d("f64", MAX_LEVELS)
    fam = FAMILIES["F1_GAUSS"]

    def timed(fn, min_passes):
        times, out, t_begin = [], None, time.perf_counter()
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            if len(times) >= min_passes and time.perf_counter() - t_begin > budget_s:
                break
        return times, out

    times, out = timed(lambda: orc.batch_adaptive1d("f64", fam, alpha, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=cores), 3)
    per_step = sum(times) / len(times)
    evals = int(out["evals"].sum())
    res = {"value": n / per_step, "unit": UNIT, "cores": cores, "kind": "port", "gevals_per_s": evals / per_step / 1e9,
           "ms_per_step": per_step * 1e3, "steps": len(times),
           "sample": f"the full workload B ({n} integrals, {evals} evaluations) x {len(times)} passes on {cores} threads (OMP_NUM_THREADS in the environment: "
                     f"{os.environ.get('OMP_NUM_THREADS', 'unset')}, overridden); scalar kernel of `quad` = oracle/fts_oracle.c ({variant}, OpenMP over the batch, IEEE, "
                     f"no reassociation); value_avx = the @turbo twin oracle/fts_oracle_avx.c (SIMD reduction + glibc libmvec exp, -ffast-math)"}
    # one core, scalar (on every 16th integral: ~0.15 s)
    one = np.ascontiguousarray(alpha[::16])
    orc.batch_adaptive1d("f64", fam, one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
    t0 = time.perf_counter()
    orc.batch_adaptive1d("f64", fam, one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
    res["value_1core"] = one.size / (time.perf_counter() - t0)
    # the @turbo twin: 256-bit (gcc's preference) and forced 512-bit vectors, the faster one is reported
    try:
        best = None
        for wide in (False, True):
            tw = OracleAvx(wide)
            tt, _ = timed(lambda: tw.batch_adaptive1d(fam, alpha, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=cores), 3)
            v = n / (sum(tt) / len(tt))
            if best is None or v > best[0]:
                t0 = time.perf_counter()
                tw.batch_adaptive1d(fam, one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
                t0 = time.perf_counter()
                tw.batch_adaptive1d(fam, one, [-1.0], [1.0], RTOL, ATOL, MAX_LEVELS, cache, nthreads=1)
                best = (v, 512 if wide else 256, one.size / (time.perf_counter() - t0))
        res["value_avx"], res["avx_vector_bits"], res["value_avx_1core"] = best
    except Exception as exc:  # pragma: no cover
        res["value_avx"], res["avx_error"] = None, str(exc)
    return res


CPU_KEYS = ("value", "value_avx", "value_1core", "value_avx_1core", "avx_vector_bits", "unit", "cores", "kind", "sample", "gevals_per_s")


# ------------------------------------------------------------------------------- helpers ------
def max_over_ranks(x: float, dev, world, dist) -> float:
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def host_link_ceiling(dev, world, dist, in_bytes: int, out_bytes: int, reps: int = 40):
    """What the box's host link gives when all N ranks move a step's bytes at once with plain pinned
    cudaMemcpyAsync (one H2D of `in_bytes` and one D2H of `out_bytes` per rank and repetition, on two
    streams): the ceiling the e2e figure is held against."""
    import torch
    src = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(out_bytes, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}
    for mode in ("both", "h2d", "d2h"):
        for timed in (False, True):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                if mode != "d2h":
                    with torch.cuda.stream(s_in):
                        d_in.copy_(src, non_blocking=True)
                if mode != "h2d":
                    with torch.cuda.stream(s_out):
                        dst.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, dev, world, dist)
        nbytes = (in_bytes if mode != "d2h" else 0) + (out_bytes if mode != "h2d" else 0)
        out[mode] = {"ms_per_step": dt / reps * 1e3, "gbs_all_ranks": nbytes * world * reps / dt / 1e9}
    return out


def sharded3d_record(fts, dev, world, rank, dist, levels, reps: int):
    """BASELINE configs[3] ("integrate3D adaptive with adaptive_cache_3D, single large integral sharded
    over outer axis"): adaptive_integrate_3D_avx (adaptive.jl:489-706) of exp(x+y+z) and
    log(1-x)log(1-y)log(1-z) on [-1,1]^3 at forced depth L (FTS_OPT_FORCE_DEPTH: with rtol = atol = 0 alone
    exp(x+y+z) stops at level 6, where two levels agree to the last bit) through the fused peer-memory exchange
    and the NCCL all-gather variant.  Strong scaling: the N = 1 arm is every rank computing the whole
    integral on its own GPU (no exchange); times are CUDA events, max over ranks, median over reps."""
    import torch
    from fts_b200 import distributed as fd
    out = []
    lo, up = [-1.0] * 3, [1.0] * 3
    for L in levels:
        cache = fts.adaptive_cache_3D(np.float64, max_levels=L)
        evals = (2 * (2 << L) + 1) ** 3
        for fam in ("exp_3d", "log_3d"):
            f = fts.builtin(fam)
            rec = {"family": fam, "level": L, "evals": evals}
            arms = [("single", dict(exchange="peer", world=1, rank=0))]
            if world > 1:
                arms += [("fused_peer", dict(exchange="peer")), ("nccl_allgather", dict(exchange="nccl"))]
            for arm, kw in arms:
                plan = fd.ShardedPlan(np.float64, f, lo, up, rtol=0.0, atol=0.0, max_levels=L, cache=cache, force_depth=True, **kw)
                ts = []
                for it in range(reps + 2):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    # the ranks leave a host barrier tens of microseconds apart, which a 0.3 ms job would count as exchange
                    # wait at its first sharded level: the streams are aligned by a device-side barrier (peer mailboxes)
                    if world > 1 and arm != "single":
                        fd.PeerMailboxes.get(None).barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    plan.launch()
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
                t = torch.tensor(ts, dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.median().item())
                s = plan.result()
                vals = torch.tensor([s[3]], dtype=torch.float64, device=dev)
                allv = [torch.zeros_like(vals) for _ in range(world)]
                if world > 1:
                    dist.all_gather(allv, vals)
                else:
                    allv = [vals]
                same = all(float(v.item()).hex() == float(allv[0].item()).hex() for v in allv)
                rec[arm] = {"ms": ms, "ms_min": float(t.min().item()), "gevals_per_s": evals / ms / 1e6, "value": float(s[3]), "level_used": int(s[6]),
                            "identical_on_all_ranks": same}
                if arm == "fused_peer":
                    w = torch.tensor(plan.exchange_wait_ns().astype(np.float64), dtype=torch.float64, device=dev)
                    wmax = w.clone()
                    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
                    rec[arm]["exchange_wait_us_per_level_rank0"] = [round(x / 1e3, 2) for x in w.cpu().tolist()]
                    rec[arm]["exchange_wait_us_per_level_max_over_ranks"] = [round(x / 1e3, 2) for x in wmax.cpu().tolist()]
                if arm in ("single", "fused_peer"):
                    # device globaltimer at each level's stop test (last repetition): where the call's time goes per level
                    done = plan.level_done_ns().astype(np.int64)
                    if done.size > 1 and done.min() > 0:
                        rec[arm]["level_us_rank0"] = [round(float(x) / 1e3, 2) for x in np.diff(done)]
            exact = (math.e - 1 / math.e) ** 3 if fam == "exp_3d" else (2 * math.log(2) - 2) ** 3
            rec["rel_err_vs_closed_form"] = abs(rec["single"]["value"] - exact) / abs(exact)
            for arm in ("fused_peer", "nccl_allgather"):
                if arm in rec:
                    rec[arm]["speedup_vs_single"] = rec["single"]["ms"] / rec[arm]["ms"]
                    rec[arm]["strong_scaling_efficiency"] = rec["single"]["ms"] / rec[arm]["ms"] / world
                    rec[arm]["rel_diff_vs_single"] = abs(rec[arm]["value"] - rec["single"]["value"]) / abs(rec["single"]["value"])
            out.append(rec)
    return out


# ------------------------------------------------------------------------------- main ---------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-log2", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--permute", action="store_true", help="shuffle the alpha sweep (seed 0) in the HEADLINE region; the default run reports it as `permuted`")
    ap.add_argument("--no-permuted", action="store_true", help="skip the secondary permuted-order measurement")
    ap.add_argument("--no-sharded3d", action="store_true", help="skip the sharded single-3D-integral record (BASELINE configs[3])")
    ap.add_argument("--sharded-levels", default="8,9", help="forced depths of the sharded 3D record")
    ap.add_argument("--no-host-link", action="store_true", help="skip the host-link ceiling probe")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the device-timed records of BASELINE configs C and E")
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
        base = cpu_reference(args.batch_log2, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
                "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config(args.batch_log2, args.gpus),
                "cpu_baseline": {k: base[k] for k in CPU_KEYS if k in base},
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

    def make_step(par):
        def step():
            fts.adaptive_batch_dev(cache, gauss, low=lo.data_ptr(), up=hi.data_ptr(), bounds_stride=0, params=par.data_ptr(), params_stride=1,
                                   rtol=RTOL, atol=ATOL, max_levels=MAX_LEVELS, batch=n, value=value.data_ptr(), err=err.data_ptr(),
                                   levels=levels.data_ptr(), flags=flags.data_ptr(), stream=stream, order="reference",
                                   options=L.OPT_QUAD_EDGE_RULES)
        return step

    step = make_step(alpha)

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

    def timed_region(step_fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        sampler.active = True
        wall0 = time.perf_counter()
        for e0, e1 in ev:
            flush.zero_()  # L2 flush, outside the event pair
            e0.record()
            step_fn()
            e1.record()
        barrier()
        wall = time.perf_counter() - wall0
        sampler.active = False
        ms = [e0.elapsed_time(e1) for e0, e1 in ev]
        return ms, max_over_ranks(sum(ms), dev, world, dist), wall

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    evals_per_step = int(fts.evals_for_level(1, levels.cpu().numpy()).sum())
    conv = int((flags & L.FLAG_CONVERGED).ne(0).sum().item())
    lv_hist = torch.bincount(levels, minlength=MAX_LEVELS + 1).cpu().tolist()

    # ---- timed region 1: inputs resident in HBM, K steps, CUDA events per step on the launch stream
    fts.launch_count(reset=True)
    step_ms, dev_ms, wall = timed_region(step, args.steps)
    launches = fts.launch_count()
    ms_per_step = dev_ms / args.steps
    total_integrals = n * world
    value_ips = total_integrals / (ms_per_step * 1e-3)
    gevals = evals_per_step * world / (ms_per_step * 1e-3) / 1e9
    ref_value = value.clone()

    # ---- secondary: the other input order of SURVEY 8(d) (shuffled unless --permute made that the headline)
    permuted = None
    if not args.no_permuted:
        _, alpha_other = workload(args.batch_log2, permute=not PERMUTE)
        other = torch.from_numpy(alpha_other).to(dev)
        step_o = make_step(other)
        for _ in range(args.warmup):
            step_o()
        p_steps = max(5, args.steps // 4)
        p_ms, p_dev, _ = timed_region(step_o, p_steps)
        permuted = {"order": "sorted sweep" if PERMUTE else "permuted, seed 0", "steps": p_steps, "ms_per_step": p_dev / p_steps,
                    "value": total_integrals / (p_dev / p_steps * 1e-3), "step_ms_min": min(p_ms), "step_ms_median": float(np.median(p_ms)),
                    "step_ms_max": max(p_ms)}

    # ---- timed region 2: end to end through the host-buffer C ABI call (pinned host memory)
    h_alpha = torch.from_numpy(alpha_np).pin_memory()
    h_lo = torch.tensor([-1.0], dtype=torch.float64).pin_memory()
    h_hi = torch.tensor([1.0], dtype=torch.float64).pin_memory()
    h_val = torch.empty(n, dtype=torch.float64).pin_memory()
    h_flg = torch.empty(n, dtype=torch.int32).pin_memory()

    def e2e_step():
        fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=h_alpha.data_ptr(),
                                params_stride=1, rtol=RTOL, atol=ATOL, max_levels=MAX_LEVELS, batch=n, value=h_val.data_ptr(),
                                flags=h_flg.data_ptr(), order="reference", options=L.OPT_QUAD_EDGE_RULES)

    e2e_steps = max(5, min(args.steps, 200))

    def e2e_region():
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        marks = [t0]
        for _ in range(e2e_steps):
            e2e_step()  # synchronous: returns when the results are in h_val / h_flg
            marks.append(time.perf_counter())
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        sampler.active = False
        return max_over_ranks(secs, dev, world, dist), np.diff(np.array(marks)) * 1e3

    # cold: the driver's W warm-up steps only (what a caller sees right after the box idled: link and
    # host power states) ...
    for _ in range(args.warmup):
        e2e_step()
    cold_s, cold_each = e2e_region()
    # ... steady state: after >= 2 s of host-link traffic
    t_warm, n_warm = time.perf_counter(), 0
    while time.perf_counter() - t_warm < 2.0 and n_warm < 20000:
        e2e_step()
        n_warm += 1
    e2e_s, e2e_each = e2e_region()
    e2e_val = total_integrals * e2e_steps / e2e_s
    e2e_ok = bool(np.array_equal(h_val.numpy(), ref_value.cpu().numpy()))
    e2e_conv = int((h_flg.numpy() & L.FLAG_CONVERGED != 0).sum())
    h2d_bytes, d2h_bytes = n * 8 + 16, n * 8 + n * 4
    link = None
    if not args.no_host_link:
        link = host_link_ceiling(dev, world, dist, h2d_bytes, d2h_bytes)

    # ---- BASELINE configs[3]: one large 3D integral sharded over the ranks (the path with an exchange step)
    sharded = None
    if not args.no_sharded3d:
        try:
            sharded = sharded3d_record(fts, dev, world, rank, dist, [int(v) for v in args.sharded_levels.split(",") if v], reps=5)
        except Exception as exc:  # pragma: no cover
            sharded = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- BASELINE configs[2] and [4] (C: 4096 2D complement boxes, FP64 and FP32; E: 65 536 Double64 quad_cmpl
    # integrals), device-resident inputs, CUDA events, rank 0: the same code as tools/bench_configs.py --cases Cdev,Edev
    other = None
    if rank == 0 and not args.no_other_configs:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs as bc
            bc.QUIET, bc.FP64_PEAK, bc.FP32_PEAK = True, peak["tflops"], fts.measure_fma_peak(np.float32, 8192)["tflops"]
            del bc.RECORDS[:]
            bc.case_cdev(10)
            bc.case_edev(5)
            other = list(bc.RECORDS)
        except Exception as exc:  # pragma: no cover
            other = {"error": f"{type(exc).__name__}: {exc}"}

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kernel_evals_per_s = evals_per_step / (ms_per_step * 1e-3)  # per GPU (max-over-ranks time)
    achieved_tflops = kernel_evals_per_s * FLOP_PER_EVAL / 1e12
    peak_tflops = peak["tflops"]
    hbm_bytes = n * (8 + 8 + 8 + 4 + 4)  # a in; value, err, level, flags out
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
    prof, prof_err = committed_profile(HEADLINE_KERNEL)
    if prof_err:
        print(f"bench.py: {prof_err}", file=sys.stderr)
    roofline = {"bound": "fp64_pipe", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                "traffic": (prof["dram_bytes_read"] + prof["dram_bytes_write"]) if prof else None,
                "peak_source": "in-run register-resident DFMA chain microbenchmark (fts_measure_fma_peak); MEASURED_PEAKS.json has no FP64 figure",
                "peak_theoretical": FP64_THEORETICAL_TFLOPS, "frac_of_theoretical": achieved_tflops / FP64_THEORETICAL_TFLOPS,
                "flop_per_eval": FLOP_PER_EVAL, "flop_per_eval_note": "cost MODEL of SURVEY 8(d) (libm-class exp = 29 flop); the shipped kernel executes fewer — see frac_executed",
                "evals_per_launch": evals_per_step, "kernel": HEADLINE_KERNEL, "kernel_ms": ms_per_step,
                "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / (ms_per_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "note": "not the binding roofline: arithmetic intensity ~290 flop/B"}}
    if prof:
        # EXECUTED arithmetic: SASS thread-instruction counts of one launch from the committed ncu capture
        # (DFMA = 2 flop, DMUL/DADD = 1), divided by that launch's evaluations, at the live evaluation rate
        ev = prof["evals_per_launch"]
        fl = (2 * prof["dfma"] + prof["dmul"] + prof["dadd"]) / ev
        roofline.update({"flop_per_eval_executed": fl, "fp64_instr_per_eval_executed": (prof["dfma"] + prof["dmul"] + prof["dadd"]) / ev,
                         "frac_executed": kernel_evals_per_s * fl / 1e12 / peak_tflops,
                         # an FP64 instruction occupies the pipe for one issue slot whatever its flop count: the
                         # pipe-occupancy fraction counts instructions at the DFMA rate (peak/2 instructions per second)
                         "frac_pipe_slots": kernel_evals_per_s * (prof["dfma"] + prof["dmul"] + prof["dadd"]) / ev * 2 / 1e12 / peak_tflops,
                         "ncu_fp64_pipe_pct_of_active": prof.get("fp64_pipe_pct_active"), "ncu_fp64_pipe_pct_of_elapsed": prof.get("fp64_pipe_pct_elapsed"),
                         "ncu_kernel_us": prof.get("time_us"), "profile": prof.get("source"), "profile_source_sha16": prof.get("source_sha16")})
    else:
        roofline["profile_error"] = prof_err
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_reference(args.batch_log2, steps=5, warmup=1, budget_s=10.0)
            cpu = {k: cpu[k] for k in CPU_KEYS if k in cpu}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
           "ms_per_step": e2e_s / e2e_steps * 1e3, "step_ms_min": float(e2e_each.min()), "step_ms_median": float(np.median(e2e_each)),
           "warmup_steps": args.warmup + e2e_steps + n_warm, "matches_device_path_bitwise": e2e_ok, "converged_from_flags": e2e_conv,
           "cold": {"value": total_integrals * e2e_steps / cold_s, "ms_per_step": cold_s / e2e_steps * 1e3, "step_ms_min": float(cold_each.min()),
                    "step_ms_median": float(np.median(cold_each)), "warmup_steps": args.warmup,
                    "note": "timed right after the driver's W warm-up steps; `value` above is the steady state after >= 2 s of host-link traffic"},
           "returns": "value[batch] (8 B) and flags[batch] (4 B: converged / max_levels, what quad's warn path needs)",
           "api": "fts_adaptive_batch on pinned host buffers: the kernel reads alpha from and writes values and flags to host memory in place over PCIe "
                  "(every step, inside the timed region); FTS_B200_ZEROCOPY=0 selects the staged 4-chunk copy pipeline"}
    if link:
        bytes_step = (h2d_bytes + d2h_bytes) * world
        e2e["host_link_peak_gbs"] = link["both"]["gbs_all_ranks"]
        e2e["achieved_host_gbs"] = bytes_step / (e2e_s / e2e_steps) / 1e9
        e2e["frac_of_host_link"] = e2e["achieved_host_gbs"] / link["both"]["gbs_all_ranks"]
        e2e["host_link"] = {"how": f"{world} rank(s) at once, per repetition one pinned cudaMemcpyAsync H2D of {h2d_bytes} B and one D2H of {d2h_bytes} B on two streams "
                                   f"(the bytes of one e2e step), 40 repetitions, wall clock, max over ranks", **link}
    line = {"metric": METRIC, "value": value_ips, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.batch_log2, world),
            "gevals_per_s": gevals, "evals_per_integral_mean": evals_per_step / n, "level_histogram": lv_hist, "converged": conv,
            "wall_s_timed_region": wall, "step_ms_min": min(step_ms), "step_ms_median": sorted(step_ms)[len(step_ms) // 2],
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "permuted": permuted, "sharded3d": sharded,
            "other_configs": other, "gpu_launches": launches, "clocks": clocks}
    if permuted:
        permuted["frac"] = evals_per_step / (permuted["ms_per_step"] * 1e-3) * FLOP_PER_EVAL / 1e12 / peak_tflops
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
