#!/usr/bin/env python
"""Fold one `ncu --set full` capture (read here, no GPU needed) into profiles/kernel_profiles.json,
the file bench.py takes `roofline.traffic` and the executed-instruction mix from.

    python tools/ncu_profile_json.py gpurun_out/prof.ncu-rep --kernel-regex k_adaptive1d_tpi --evals 243758272 \
        --files fts_dd.cuh,fts_math.cuh,fts_exp_table.inc,fts_families.cuh,fts_kernels.cuh --note "bench.py --steps 2"

One record per kernel NAME (the last matching launch of the capture wins); every record carries the
hash of the device sources it was taken at, and bench.py / tests refuse a record whose hash no
longer matches the tree.
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M = {"time_us": "gpu__time_duration.sum", "dfma": "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
     "dmul": "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dadd": "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
     "ffma": "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "warp_inst": "smsp__inst_executed.sum",
     "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
     "fp64_pipe_pct_active": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "fp64_pipe_pct_elapsed": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
     "grid": "launch__grid_size", "block": "launch__block_size", "waves_per_sm": "launch__waves_per_multiprocessor",
     "eligible_warps": "smsp__warps_eligible.avg.per_cycle_active",
     "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "msecond": 1e3, "ms": 1e3, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}


def num(v, unit):
    v = float(v.replace(",", ""))
    return v * SCALE.get(unit, 1.0)


def short_name(full: str) -> str:
    """void fts::k_x<double, fts::F1Gauss, (bool)0>(fts::Args<double>) -> k_x<double, F1Gauss, 0>  (ncu prints booleans as 0/1)"""
    s = re.sub(r"^void\s+", "", full.strip())
    if s.endswith(")"):  # strip the parameter list: the '(' that matches the last ')'
        depth = 0
        for i in range(len(s) - 1, -1, -1):
            if s[i] == ")":
                depth += 1
            elif s[i] == "(":
                depth -= 1
                if depth == 0:
                    s = s[:i]
                    break
    s = s.replace("fts::", "").replace("(bool)0", "0").replace("(bool)1", "1").replace("false", "0").replace("true", "1")
    s = re.sub(r"\(int\)(\d+)", r"\1", s)
    return s.strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel-regex", default=".")
    ap.add_argument("--evals", type=int, default=0, help="integrand evaluations of ONE launch of the kernel (bench line: roofline.evals_per_launch)")
    ap.add_argument("--files", default="fts_dd.cuh,fts_math.cuh,fts_exp_table.inc,fts_families.cuh,fts_kernels.cuh")
    ap.add_argument("--note", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "kernel_profiles.json"))
    a = ap.parse_args()
    import bench
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    db = json.load(open(a.out)) if os.path.exists(a.out) else {}
    files = tuple(f for f in a.files.split(",") if f)
    for r in rows[2:]:
        full = r[hdr.index("Kernel Name")]
        if not re.search(a.kernel_regex, full):
            continue
        rec = {"kernel_full": full, "source": os.path.basename(a.rep), "note": a.note, "source_files": list(files), "source_sha16": bench.kernel_source_sha(files),
               "evals_per_launch": a.evals}
        cyc = float(r[hdr.index("smsp__cycles_elapsed.avg")].replace(",", "")) if "smsp__cycles_elapsed.avg" in hdr else None
        for k, m in M.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    rec[k] = num(r[i], units[i])
                except ValueError:
                    rec[k] = None
            elif m + ".per_cycle_elapsed" in hdr and cyc:
                # `--set full` of this ncu reports the SASS op counters as <sum>.per_cycle_elapsed: count = rate x elapsed cycles
                rec[k] = float(r[hdr.index(m + ".per_cycle_elapsed")].replace(",", "")) * cyc
        if a.evals:
            rec["fp64_instr_per_eval"] = (rec["dfma"] + rec["dmul"] + rec["dadd"]) / a.evals
            rec["warp_inst_per_eval"] = rec["warp_inst"] * 32 / a.evals
        db[short_name(full)] = rec
        print(short_name(full), json.dumps({k: rec.get(k) for k in ("time_us", "fp64_instr_per_eval", "fp64_pipe_pct_active", "dram_bytes_read", "registers")}))
    json.dump(db, open(a.out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
