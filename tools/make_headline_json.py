#!/usr/bin/env python
"""profiles/headline_ncu.json (what bench.py prints under roofline.ncu / uses for roofline.traffic) from a condensed ncu
summary: python tools/make_headline_json.py profiles/<summary>.csv <kernel> <paths> <nsteps> "<capture description>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kernel, paths, nsteps, desc = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1, "us": 1e-3, "s": 1e3}
m = {}
for row in csv.reader(open(path)):
    if len(row) >= 5 and row[1] == kernel:
        try:
            m[row[2]] = float(row[4]) * unit.get(row[3], 1)
        except ValueError:
            pass
out = {
    "kernel": kernel, "capture": desc, "file": os.path.relpath(path, ROOT), "paths": paths, "nsteps": nsteps,
    "duration_ms": m["gpu__time_duration.sum"],
    "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]),
    "algorithmic_bytes": 16 * paths,
    "fp64_pipe_active_pct": round(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], 2),
    "issue_active_pct": round(m["smsp__issue_active.avg.pct_of_peak_sustained_active"], 2),
    "registers_per_thread": int(m["launch__registers_per_thread"]),
    "warps_active_per_sm": round(m["sm__warps_active.avg.per_cycle_active"], 2),
}
json.dump(out, open(os.path.join(ROOT, "profiles", "headline_ncu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
