#!/usr/bin/env python
"""All five BASELINE.json configurations on one GPU, with the CPU restatement timed beside them on a bounded sample.
Writes gpurun_out/configs.json (copied to profiles/ by hand).  C2 is bench.py's headline; here it is repeated once."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkg
import oracle as O
import torch

S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
HP, BP = (0.01, 10.0, 0.3, 0.1, 0.4), (0.01, 0.3)
out = {}
hbm = 6547.8
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass
peak64, _ = S.measure_peak("f64")
peak32, _ = S.measure_peak("f32")
out["peaks"] = {"fp64_tflops_measured": peak64, "fp32_tflops_measured": peak32, "hbm_gbs_measured": hbm}


def best_ms(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        fn()
        t.append(S.last_kernel_ms())
    return min(t)


# C1: BlackScholes EM 10^6 x 252, terminal mean
m, sch = S.BlackScholes(*BP), S.EulerMaruyama(1 / 252)
e = S.estimate(m, sch, S.Moments(), 0.0, 100.0, 252, 10 ** 6, seed=0)
ms = best_ms(lambda: S.estimate(m, sch, S.Moments(), 0.0, 100.0, 252, 10 ** 6, seed=0))
t = time.perf_counter()
sums, cores, _ = O.sample_rng("BlackScholes", 0, BP, 0.0, 1 / 252, 100.0, 252, 10 ** 6, 0, 0)
cpu_s = time.perf_counter() - t
out["C1"] = {"workload": "BlackScholes EM, 10^6 x 252, terminal mean", "gpu_ms": ms, "path_steps_per_s": 2.52e8 / ms * 1e3,
             "mean": float(e.mean[0]), "stderr": float(e.stderr[0]), "closed_form": 101.00499666826876,
             "cpu_path_steps_per_s": 2.52e8 / cpu_s, "cpu_cores": cores, "cpu_mean": sums[0] / 1e6}

# C2 / C5: Heston EM, FP64 and FP32 sweep
m, sch = S.Heston(*HP), S.EulerMaruyama(1 / 1024)
sweep = []
for prec in ("f64", "f32"):
    for n in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 10 ** 9):
        if n == 10 ** 9 and prec == "f64" and "--big" not in sys.argv:
            continue
        e = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], 1024, n, seed=1, precision=prec)
        ms = S.last_kernel_ms()
        if n <= 10 ** 7:
            ms = best_ms(lambda: S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], 1024, n, seed=1, precision=prec))
        zS = (e.mean[0] - 101.00501177656254) / e.stderr[0]
        zV = (e.mean[1] - 0.3000043222543304) / e.stderr[1]
        sweep.append({"precision": prec, "npaths": n, "ms": ms, "path_steps_per_s": n * 1024 / ms * 1e3,
                      "mean_S": float(e.mean[0]), "stderr_S": float(e.stderr[0]), "z_S": float(zS),
                      "mean_V": float(e.mean[1]), "z_V": float(zV), "nonfinite": e.nonfinite})
        print(sweep[-1], flush=True)
out["C2_C5_heston_sweep"] = sweep
t = time.perf_counter()
sums, cores, _ = O.sample_rng("Heston", 0, HP, 0.0, 1 / 1024, [100.0, 0.4], 1024, 2 * 10 ** 6, 0, 0)
cpu_s = time.perf_counter() - t
out["C2_cpu"] = {"path_steps_per_s": 2e6 * 1024 / cpu_s, "cores": cores, "sample": "2*10^6 paths x 1024 steps"}

# C3: BlackScholes Milstein full paths 10^7 x 512 (41 GB FP64 / 20.5 GB FP32, device resident)
m, sch = S.BlackScholes(*BP), S.Milstein(1 / 512)
c3 = []
for prec, tdt in (("f64", torch.float64), ("f32", torch.float32)):
    for lay in ("soa", "reference"):
        n = 10 ** 7
        shape = (513, n) if lay == "soa" else (n, 513)
        buf = torch.empty(shape, dtype=tdt, device="cuda")
        ms = best_ms(lambda: S.simulate(m, sch, 0.0, 100.0, 512, n, seed=2, precision=prec, out=buf, layout=lay), reps=2)
        nbytes = buf.numel() * buf.element_size()
        last = (buf[-1] if lay == "soa" else buf[:, -1]).double()
        c3.append({"precision": prec, "layout": lay, "ms": ms, "path_steps_per_s": n * 512 / ms * 1e3, "bytes": nbytes,
                   "write_GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / hbm,
                   "mean_S_T": float(last.mean()), "closed_form": 101.00500684477365,
                   "stderr": float(last.std() / math.sqrt(n))})
        print(c3[-1], flush=True)
        del buf
        torch.cuda.empty_cache()
out["C3_bs_milstein_fullpath"] = c3
t = time.perf_counter()
x, cores = O.simulate_rng("BlackScholes", 1, BP, 0.0, 1 / 512, 100.0, 512, 10 ** 5, 0, 0)
cpu_s = time.perf_counter() - t
out["C3_cpu"] = {"path_steps_per_s": 1e5 * 512 / cpu_s, "cores": cores, "sample": "10^5 paths x 512 steps"}

# C4: Heston MLMC European call, levels 0..8, base dt 1/16, substeps 2
ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, range(0, 9))
m = S.Heston(*HP)
pay = S.Call(100.0, math.exp(-0.01))
nl = [2 ** (26 - 2 * l) for l in range(9)]
est = S.ml_estimate(m, ml, pay, 0.0, [100.0, 0.4], 16, nl, seed=3)
ms = best_ms(lambda: S.ml_estimate(m, ml, pay, 0.0, [100.0, 0.4], 16, nl, seed=3), reps=2)
fine_steps = sum(n * 16 * 2 ** l for l, n in enumerate(nl))
coarse_steps = sum(n * 16 * 2 ** (l - 1) for l, n in enumerate(nl) if l > 0)
t = time.perf_counter()
cpu_steps = 0
for l in range(1, 9):
    n = max(nl[l] // 100, 8)
    O.multilevel_rng("Heston", 0, HP, 0.0, (1 / 16) / 2 ** (l - 1), (1 / 16) / 2 ** l, 2, [100.0, 0.4], 16 * 2 ** (l - 1), n, 0, 0)
    cpu_steps += n * 16 * 2 ** l
cpu_s = time.perf_counter() - t
out["C4_heston_mlmc_call"] = {
    "levels": "0:8", "npaths": nl, "ms": ms, "fine_path_steps": fine_steps, "coarse_path_steps": coarse_steps,
    "fine_path_steps_per_s": fine_steps / ms * 1e3, "estimate": est.mean, "stderr": est.stderr,
    "semi_analytic_continuous_time": 22.3374853590066, "level_mean": est.level_mean.tolist(),
    "level_var": est.level_var.tolist(), "nonfinite": est.nonfinite.tolist(),
    "cpu_fine_path_steps_per_s": cpu_steps / cpu_s, "cpu_cores": O.max_threads(), "cpu_sample": "levels 1..8 with 1/100 of the paths"}
print(out["C4_heston_mlmc_call"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out["normal_stream"] = S.simulation.DEFAULT_RNG or 1
name = "configs_v%d.json" % (S.simulation.DEFAULT_RNG or 1)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
print("wrote gpurun_out/" + name)
