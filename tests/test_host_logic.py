"""Host-side logic: scheme types, MultilevelScheme helpers, the budget allocator (against the oracle's
line-by-line restatement), sharding, and the multi-rank combination of exact sums over gloo (world size 2,
CPU).  Runs without a GPU."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scheme_types(S):
    # schemes.jl:3-11,51-54
    e = S.EulerMaruyama(0.1)
    assert e.dt == 0.1 and e.Δt == 0.1
    assert S.EulerMaruyama(0.1, 0.0, 1.0) == e          # T(Δt, t, s) = T(Δt)
    f = S.subdivide(e, 4)
    assert isinstance(f, S.EulerMaruyama) and f.dt == 0.1 / 4
    assert isinstance(S.subdivide(S.Milstein(0.5), 2), S.Milstein)
    assert S.subdivide(S.RungeKutta(0.5), 5) == S.RungeKutta(0.1)


@pytest.mark.parametrize("substeps,lo,hi,budget", [(4, 0, 4, 1000), (2, 0, 8, 2 ** 20), (2, 0, 8, 2 ** 34), (3, 1, 5, 12345.6),
                                                   (2, 0, 0, 17), (4, 2, 3, 10 ** 6)])
def test_npaths_and_cost_match_the_oracle(S, O, substeps, lo, hi, budget):
    ml = S.MultilevelScheme(S.EulerMaruyama(0.1), substeps, range(lo, hi + 1))
    n = S.npaths(ml, budget)
    assert list(n) == list(O.npaths(substeps, lo, hi, budget))
    assert list(S.cost(ml, n)) == list(O.cost(substeps, lo, hi, n))
    assert [s.dt for s in S.schemes(ml)] == [O.level_dt(0.1, substeps, l) for l in range(lo, hi + 1)]


def test_shard_ranges_tile_the_run(S):
    for n in (0, 1, 1000, 1024, 10 ** 6 + 7, 10 ** 8):
        for world in (1, 2, 3, 4, 8):
            r = [S.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(lo % S._lib.SHARD_ALIGN == 0 for lo, hi in r if lo < n)
            if n >= world * 8 * S._lib.SHARD_ALIGN:
                sizes = [hi - lo for lo, hi in r]
                assert max(sizes) - min(sizes) <= 2 * S._lib.SHARD_ALIGN


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import _pkg
    import oracle as O
    S = _pkg.load()
    L = S._lib.lib()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = S.shard_range(n, rank, world, align=64)
    # this rank's per-chunk partial sums (the oracle plays the kernel: test infrastructure)
    z = O.normals_paths(3, 0, lo, hi - lo, 8, 2)
    x = O.sample_z("Heston", 0, (0.01, 10.0, 0.3, 0.1, 0.4), 0.0, 1 / 8, [100.0, 0.4], z)
    bins = np.zeros((2, S._lib.NBINS + 1), dtype=np.int64)
    P64 = C.POINTER(C.c_int64)
    for c0 in range(0, hi - lo, 64):          # chunk partials, then exact accumulation
        part = x[c0:c0 + 64]
        L.sdeb200_bins_add(bins[0, :-1].ctypes.data_as(P64), float(np.sum(part[:, 0])))
        L.sdeb200_bins_add(bins[1, :-1].ctypes.data_as(P64), float(np.sum(part[:, 0] ** 2)))
    bins[:, -1] = hi - lo
    t = torch.from_numpy(bins)
    dist.all_reduce(t)                         # what ncclAllReduce(int64, sum) does on the GPUs
    out = [L.sdeb200_bins_to_double(np.ascontiguousarray(bins[k, :-1]).ctypes.data_as(P64)) for k in range(2)]
    if rank == 0:
        q.put((out, int(bins[0, -1])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_multi_rank_exact_reduction_over_gloo(S, O, world):
    """Two CPU ranks (gloo) shard one logical run, all-reduce the integer limbs, and obtain bit-identical sums
    to the single-rank run — the N > 1 path of the estimators, minus the kernels."""
    import torch.multiprocessing as mp
    n, port = 64 * 37, 29500 + os.getpid() % 1000 + world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, cnt = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference of the same chunked computation
    L = S._lib.lib()
    z = O.normals_paths(3, 0, 0, n, 8, 2)
    x = O.sample_z("Heston", 0, (0.01, 10.0, 0.3, 0.1, 0.4), 0.0, 1 / 8, [100.0, 0.4], z)
    P64 = C.POINTER(C.c_int64)
    b = np.zeros((2, S._lib.NBINS), dtype=np.int64)
    for c0 in range(0, n, 64):
        L.sdeb200_bins_add(b[0].ctypes.data_as(P64), float(np.sum(x[c0:c0 + 64, 0])))
        L.sdeb200_bins_add(b[1].ctypes.data_as(P64), float(np.sum(x[c0:c0 + 64, 0] ** 2)))
    want = [L.sdeb200_bins_to_double(b[k].ctypes.data_as(P64)) for k in range(2)]
    assert cnt == n and got == want
    assert S.distributed.combine_bins_host([b[0], np.zeros_like(b[0])]) == want[0]
