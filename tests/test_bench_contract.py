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


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_only_rank0_prints():
    """N > 1: rank 0 alone runs and prints the reference line; the other ranks exit 0 without work."""
    env = {"WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29433"}
    l0 = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch-log2", "12"], dict(env, RANK="0", LOCAL_RANK="0"))
    l1 = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch-log2", "12"], dict(env, RANK="1", LOCAL_RANK="1"))
    assert json.loads(l0[-1])["impl"] == "reference" and json.loads(l0[-1])["n_gpus"] == 2
    assert l1 == [] or all(not s.startswith("{") for s in l1)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_our_arm_line():
    lines = _run(["--steps", "5", "--warmup", "3", "--no-cpu-baseline"])
    d = json.loads(lines[-1])
    assert COMMON <= set(d) and {"roofline", "clocks"} <= set(d)
    r, e = d["roofline"], d["e2e"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.001 and e["matches_device_path_bitwise"]
    assert d["gpu_launches"] >= d["steps"] and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
