"""bench.py contract (the driver depends on it): one JSON line per arm with the agreed keys.
The reference arm runs on the CPU and is checked here; our arm needs a B200 (`-m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
          "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


@pytest.mark.timeout(600)
def test_reference_arm_line():
    lines = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--batch-log2", "14"])
    d = json.loads(lines[-1])
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "batched_adaptive_integrals_per_s_fp64" and d["unit"] == "integrals/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value_1core"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the @turbo twin (adaptive.jl:94-133) is timed beside the scalar kernel `quad` runs
    assert d["cpu_baseline"]["value_avx"] > 0 and d["cpu_baseline"]["avx_vector_bits"] in (256, 512)
    # the whole workload is timed (no extrapolation from a sub-sample)
    assert "full workload" in d["cpu_baseline"]["sample"]


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_only_rank0_prints():
    """N > 1: rank 0 alone runs and prints the reference line; the other ranks exit 0 without work."""
    env = {"WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29433"}
    l0 = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch-log2", "12"], dict(env, RANK="0", LOCAL_RANK="0"))
    l1 = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch-log2", "12"], dict(env, RANK="1", LOCAL_RANK="1"))
    d0 = json.loads(l0[-1])
    assert d0["impl"] == "reference" and d0["n_gpus"] == 2
    assert l1 == [] or all(not s.startswith("{") for s in l1)


@pytest.mark.timeout(600)
def test_reference_arm_uses_all_cores_under_torchrun_environment():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use every
    core of its affinity mask (round 1 ran it on ONE core at N > 1 and inflated the ratios 20x), and
    its `config` must be the one our arm prints (the driver's same_config check)."""
    import importlib.util
    env = {"WORLD_SIZE": "2", "RANK": "0", "LOCAL_RANK": "0", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29434", "OMP_NUM_THREADS": "1"}
    d = json.loads(_run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--batch-log2", "14"], env)[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    if d["cpu_baseline"]["cores"] >= 4:
        assert d["cpu_baseline"]["value"] > 1.5 * d["cpu_baseline"]["value_1core"]
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert d["config"] == mod.config(14, 2)


def test_committed_kernel_profiles_match_the_kernel_sources():
    """bench.py reads roofline.traffic and the executed-instruction mix from the committed ncu capture
    (profiles/kernel_profiles.json) and refuses a capture taken at other kernel sources: editing a
    kernel header without re-profiling fails HERE, loudly, not in a silently stale bench line."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rec, err = mod.committed_profile(mod.HEADLINE_KERNEL)
    assert err is None, err
    assert rec["evals_per_launch"] > 0 and rec["dfma"] > 0 and rec["dram_bytes_read"] > 0 and rec["source_sha16"] == mod.kernel_source_sha(tuple(rec["source_files"]))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_our_arm_line():
    lines = _run(["--steps", "5", "--warmup", "3", "--no-cpu-baseline"])
    d = json.loads(lines[-1])
    assert COMMON <= set(d) and {"roofline", "clocks"} <= set(d)
    r, e = d["roofline"], d["e2e"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and 0 < r["frac_executed"] <= r["frac"] * 1.05 and 10 < r["fp64_instr_per_eval_executed"] < 25
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.001 and e["matches_device_path_bitwise"]
    assert e["d2h_bytes_per_step"] == (1 << 20) * 12 and e["converged_from_flags"] == 1 << 20       # values AND flags come back
    assert e["cold"]["value"] > 0 and 0 < e["frac_of_host_link"] < 3.0 and e["host_link_peak_gbs"] > 10
    assert d["permuted"]["value"] > 0 and 0 < d["permuted"]["frac"] < 1.2
    sh = d["sharded3d"]
    assert isinstance(sh, list) and {(x["family"], x["level"]) for x in sh} == {("exp_3d", 8), ("log_3d", 8), ("exp_3d", 9), ("log_3d", 9)}
    assert all(x["single"]["level_used"] == x["level"] and x["rel_err_vs_closed_form"] < 1e-11 for x in sh)
    oc = d["other_configs"]
    assert isinstance(oc, list) and {(x["case"], x.get("dtype") or x.get("family")) for x in oc} == {
        ("C-device", "float64"), ("C-device", "float32"), ("E-device", "c_invsqrt"), ("E-device", "c_log_bmx")}
    assert all(x["ms"] > 0 and x["gevals_per_s"] > 0 for x in oc) and all(x["max_rel_err_vs_exact"] < 1e-5 for x in oc if x["case"] == "C-device")
    assert d["gpu_launches"] >= d["steps"] and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
