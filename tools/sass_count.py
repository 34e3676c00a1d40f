#!/usr/bin/env python
"""Count the instructions of the hot loop of the headline kernels from SASS (cuobjdump, no GPU needed) and write
profiles/sass_counts.json.  bench.py uses it for the `executed` roofline figure:
FP flops per path-step as executed = (2*FMA + MUL + ADD) of the main loop / path-steps per loop iteration."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "sdemodels.jl_b200", "csrc", "build", "sde_aot_fast.o")
PPT = 4
# kernel -> path-steps per iteration of its main loop (PPT paths x U steps per Philox batch)
KERNELS = {
    "heston_sample_em_f64": ("sdek_heston_f_sample_em_f64", PPT * 1, "D"),
    "heston_sample_em_f64v2": ("sdek_heston_f_sample_em_f64v2", PPT * 4, "D"),   # SDE_PPT_V2 = 4 paths x 4 steps per batch
    "heston_sample_em_f32": ("sdek_heston_f_sample_em_f32", PPT * 2, "F"),
    "bs_sample_em_f64": ("sdek_bs_f_sample_em_f64", PPT * 2, "D"),
    "bs_simsoa_mil_f64": ("sdek_bs_f_simsoa_mil_f64", 4 * 2, "D"),
    "bs_simsoa_mil_f32": ("sdek_bs_f_simsoa_mil_f32", 4 * 4, "F"),
    "heston_mlmc_em_f64": ("sdek_heston_f_mlmc_em_f64", 2, "D"),  # substeps == 2 path: two fine steps per iteration
}


def loop_counts(fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, OBJ], capture_output=True, text=True, check=True).stdout
    ins = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    # innermost loops only (the kernels wrap the time loop in a chunk loop); the hot one is the largest of those
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    # the time loop is the first innermost loop that contains a full Philox block (out-of-line helpers, if any, come
    # later in the listing and may have loops of their own)
    def philox(l):
        return sum(1 for a, t in ins if l[0] <= a <= l[1] and "IMAD.WIDE" in t)
    cands = [l for l in sorted(inner) if philox(l) >= 9]
    best = cands[0] if cands else max(inner, key=lambda l: l[1] - l[0])
    cnt = collections.Counter()
    for a, t in ins:
        if best[0] <= a <= best[1]:
            t = re.sub(r"^@!?U?P\d\s+", "", t)
            cnt[t.split()[0].split(".")[0]] += 1
    return cnt


def main():
    out = {}
    for key, (fun, steps, ty) in KERNELS.items():
        c = loop_counts(fun)
        total = sum(c.values())
        fma, mul, add = c[ty + "FMA"], c[ty + "MUL"], c[ty + "ADD"]
        fp = fma + mul + add + c[ty + "SETP"] + c[ty + "MNMX"]
        out[key] = {
            "kernel": fun, "path_steps_per_loop_iteration": steps, "loop_instructions": total,
            "instructions_per_path_step": total / steps, "fp_instructions_per_path_step": fp / steps,
            "fp_flops_per_path_step": (2 * fma + mul + add) / steps,
            "mix": dict(c.most_common()),
        }
        print(f"{key:24s} {total / steps:7.1f} instr/path-step, {fp / steps:6.1f} FP, {(2 * fma + mul + add) / steps:6.1f} flops")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "sass_counts.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    sys.exit(main())
