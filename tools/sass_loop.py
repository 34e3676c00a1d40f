#!/usr/bin/env python
"""Print the hot loop of a kernel from SASS with a full-mnemonic histogram: python tools/sass_loop.py <kernel> [obj] [-l]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fun = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else os.path.join(ROOT, "sdemodels.jl_b200/csrc/build/sde_aot_fast.o")
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
ins = []
for l in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
def philox(l):
    return sum(1 for a, t in ins if l[0] <= a <= l[1] and "IMAD.WIDE" in t)
cands = [l for l in sorted(inner) if philox(l) >= 9]
best = cands[0] if cands else max(inner, key=lambda l: l[1] - l[0])
cnt = collections.Counter()
body = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
for a, t in body:
    t2 = re.sub(r"^@!?U?P\d\s+", "", t)
    cnt[t2.split()[0]] += 1
print(f"loop 0x{best[0]:x}..0x{best[1]:x}: {len(body)} instructions")
for k, v in cnt.most_common():
    print(f"  {v:5d} {k}")
if "-l" in sys.argv:
    for a, t in body:
        print(f"{a:05x} {t}")
