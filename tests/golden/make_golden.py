"""Extract the reference's published result values (the golden vectors of SURVEY.md §8c).

Reads /root/reference/benchmark/results/timings*.csv (FastTanhSinhQuadrature.jl v2.1.0) and keeps
the rows produced by the reference's own kernels ("FTS adaptive (typed)" = adaptive_integrate_*D,
"FTS quad (convenience)" = quad, "FTS avx (precomputed)" = integrate*D_avx at N = 2^(level+2)),
with the settings they were produced with (benchmark/benchmarks.jl:11-20: rtol=1e-6, atol=1e-8,
max_levels 12/9/7).  Run in the build container only (the reference is absent on the GPU box):

    python tests/golden/make_golden.py > tests/golden/reference_csv_fts.json
"""
import csv
import json
import re
import sys

REF = "/root/reference/benchmark/results"
FILES = ["timings_julia1.12.csv", "timings_julia1.13.csv", "timings.csv"]
METHODS = {"FTS adaptive (typed)": "adaptive", "FTS quad (convenience)": "quad",
           "FTS avx (precomputed)": "avx"}

rows = []
for fn in FILES:
    with open(f"{REF}/{fn}") as fh:
        for line_no, r in enumerate(csv.DictReader(fh), start=2):
            if r["method"] not in METHODS:
                continue
            m = re.search(r"N=(\d+); adaptive_level=(\d+)", r["notes"] or "")
            rows.append(dict(
                source=f"benchmark/results/{fn}:{line_no}", dim=int(r["dim"][0]), function=r["function"],
                kind=METHODS[r["method"]], value=r["value"], exact=r["exact"],
                N=int(m.group(1)) if m else None, adaptive_level=int(m.group(2)) if m else None,
                max_levels_reached="max_levels_reached" in (r["notes"] or "")))
json.dump(dict(settings=dict(rtol=1e-6, atol=1e-8, max_levels={"1": 12, "2": 9, "3": 7},
                             source="benchmark/benchmarks.jl:11-20"), rows=rows), sys.stdout, indent=1)
