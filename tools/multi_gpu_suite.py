#!/usr/bin/env python
"""Multi-GPU measurements beyond bench.py's weak / strong headline lines:
   C4 (Heston MLMC call, levels 0:8) and C5's small points (10^5 .. 10^7 paths IN TOTAL, strong scaling) at N GPUs, with the
   pieces timed separately so that the limiter of the small cases is MEASURED: kernel-only (no communicator), reduction-only
   (an estimator call with zero paths: memset + all-reduce + read-back) and the whole call.

   torchrun --nproc-per-node N tools/multi_gpu_suite.py         one process per GPU, one NCCL all-reduce per estimator / level
   python tools/multi_gpu_suite.py --single N                    ONE process, N GPUs (sdeb200_init_devices): worker thread per
                                                                 device, exact sums combined on the host
Writes gpurun_out/multigpu_<mode>_N<N>.json."""
import json
import math
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg

S = _pkg.load()
single = "--single" in sys.argv
HP = (0.01, 10.0, 0.3, 0.1, 0.4)
he, sch = S.Heston(*HP), S.EulerMaruyama(1 / 1024)
X0 = [100.0, 0.4]
RNG = 2

if single:
    N = int(sys.argv[sys.argv.index("--single") + 1])
    rank, world = 0, 1
    nd = S.init_devices(N)
    assert nd == N, (nd, N)
    dist = None
else:
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    N = world
    S._lib.check(S._lib.lib().sdeb200_init(local))
    import torch
    torch.cuda.set_device(local)
    dist = None


def barrier():
    if dist is not None:
        dist.barrier()


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        barrier()
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    ms = statistics.median(ts)
    if dist is not None:
        import torch
        v = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms = float(v.item())
    return ms


out = {"mode": "one process, %d GPUs" % N if single else "one process per GPU, %d ranks" % N, "n_gpus": N, "normal_stream": RNG}
sizes = [10 ** 5, 10 ** 6, 10 ** 7]


def share(n):
    return (0, n) if single else S.shard_range(n, rank, world)


# (1) kernel-only: this rank's share with no communicator (multi-process) — what the GPU work alone costs
kernel_only = {}
if not single:
    for n in sizes:
        lo, hi = share(n)
        kernel_only[n] = timed(lambda: S.estimate(he, sch, S.Moments(), 0.0, X0, 1024, hi - lo, seed=1, path_offset=lo, rng=RNG), 10)
    S.init_distributed(local)
    import torch.distributed as dist   # noqa: F811
# (2) reduction-only: zero paths -> memset + (all-reduce) + read-back + rounding
red_only = timed(lambda: S.estimate(he, sch, S.Moments(), 0.0, X0, 1024, 0, seed=1, rng=RNG), 30)
# (3) the whole call, strong scaling (n paths IN TOTAL)
c5 = []
for n in sizes + [10 ** 8]:
    lo, hi = share(n)
    reps = 10 if n <= 10 ** 7 else 3
    ms = timed(lambda: S.estimate(he, sch, S.Moments(), 0.0, X0, 1024, hi - lo, seed=1, path_offset=lo, rng=RNG), reps)
    e = S.estimate(he, sch, S.Moments(), 0.0, X0, 1024, hi - lo, seed=1, path_offset=lo, rng=RNG)
    c5.append({"npaths_total": n, "ms": ms, "path_steps_per_s": n * 1024 / ms * 1e3, "kernel_only_ms": kernel_only.get(n),
               "reduction_only_ms": red_only, "mean_S": float(e.mean[0]), "n": int(e.n),
               "z_S": float((e.mean[0] - 101.00501177656254) / e.stderr[0])})
out["C5_strong"] = c5
# C4: levels 0:8, N_l = 2^(26-2l) in total, every level sharded over all GPUs, one reduction per level
ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, range(0, 9))
pay = S.Call(100.0, math.exp(-0.01))
nl = [2 ** (26 - 2 * l) for l in range(9)]
if single:
    counts, offs = nl, None
else:
    rngs = [S.shard_range(k, rank, world) for k in nl]
    counts, offs = [b - a for a, b in rngs], [a for a, b in rngs]
ms = timed(lambda: S.ml_estimate(he, ml, pay, 0.0, X0, 16, counts, seed=3, rng=RNG, level_offsets=offs), 5)
est = S.ml_estimate(he, ml, pay, 0.0, X0, 16, counts, seed=3, rng=RNG, level_offsets=offs)
fine = sum(n * 16 * 2 ** l for l, n in enumerate(nl))
out["C4"] = {"ms": ms, "fine_path_steps_per_s": fine / ms * 1e3, "estimate": est.mean, "stderr": est.stderr, "n": est.n.tolist(),
             "level_sums_Y": est.sums[:, 0].tolist()}
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = "multigpu_%s_N%d.json" % ("single" if single else "ranks", N)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
    print(json.dumps(out)[:1500])
if dist is not None:
    dist.barrier()
    S._lib.lib().sdeb200_comm_destroy()
    dist.destroy_process_group()
