#!/usr/bin/env python
"""Condense .ncu-rep captures (gpurun_out/) into the long-format CSV kept under profiles/:
   python tools/ncu_summary.py out.csv capture_name=report.ncu-rep [...]
Keeps the metrics the roofline discussion uses (duration, DRAM / L2 bytes, pipe utilisation, issue slots, occupancy,
stall reasons); ncu itself is run here on the CPU box in import mode (no GPU needed)."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sectors_srcunit_tex_op_(read|write)\.sum|lts__t_bytes\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"launch__(registers_per_thread|occupancy_limit_\w+|shared_mem_per_block_(dynamic|static)|grid_size|block_size)|"
    r"sm__inst_executed_pipe_(fp64|fma|alu|xu|lsu|uniform)\.avg\.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.per_cycle_active|smsp__inst_executed\.sum|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__compute_memory_throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def main():
    out = sys.argv[1]
    rows = [("capture", "kernel", "metric", "unit", "value")]
    for spec in sys.argv[2:]:
        name, path = spec.split("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        r = list(csv.reader(io.StringIO(txt)))
        hdr, units = r[0], r[1]
        kcol = hdr.index("Kernel Name")
        for line in r[2:]:
            if len(line) < len(hdr):
                continue
            for i, h in enumerate(hdr):
                if KEEP.match(h):
                    rows.append((name, line[kcol], h, units[i], line[i]))
    with open(out, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    print("wrote", out, len(rows) - 1, "rows")


if __name__ == "__main__":
    main()
