"""One process, several GPUs (sdeb200_init_devices): host-memory calls are sharded over the bound devices —
contiguous path ranges on SDEB200_SHARD_ALIGN boundaries, one worker thread per device, the estimators' exact integer
sums added on the host — and must return THE SAME BITS as one GPU.  On a single-GPU box the test hook
SDEB200_FAKE_DEVICES binds several logical devices to the one physical GPU, which exercises the same sharding, threading
and combination code; on a multi-GPU box the real devices are used.  Also here: pageable vs pinned host outputs, the
library-owned device buffers, non-finite reporting of the trajectory drivers, the level-stream namespace."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import BS_P, HESTON_P

pytestmark = pytest.mark.gpu


@pytest.fixture()
def multi(gpu):
    S = gpu
    n = C.c_int()
    S._lib.lib().sdeb200_device_count(C.byref(n))
    fake = n.value < 2
    if fake:
        os.environ["SDEB200_FAKE_DEVICES"] = "3"
    try:
        nd = S.init_devices(0 if not fake else 3)
        assert nd >= 2 and S.device_info() == (nd, 0)
        yield S, nd
    finally:
        os.environ.pop("SDEB200_FAKE_DEVICES", None)
        S._lib.check(S._lib.lib().sdeb200_init(0))
        assert S.device_info() == (1, 0)


def _single(S, fn):
    S._lib.check(S._lib.lib().sdeb200_init(0))
    return fn()


def test_sharded_calls_equal_one_gpu_bit_for_bit(multi, O):
    S, nd = multi
    he, sch = S.Heston(*HESTON_P), S.EulerMaruyama(1 / 64)
    bs = S.BlackScholes(*BS_P)
    n = 50_000 + 37                      # not a multiple of the shard alignment
    runs = {
        "sample": lambda: S.sample(he, sch, 0.0, [100.0, 0.4], 64, n, seed=5),
        "sample_v2": lambda: S.sample(he, sch, 0.0, [100.0, 0.4], 64, n, seed=5, rng=2),
        "simulate": lambda: S.simulate(bs, S.Milstein(1 / 32), 0.0, 100.0, 32, 9000, seed=6)[0],
        "estimate": lambda: S.estimate(he, sch, S.Moments(), 0.0, [100.0, 0.4], 64, n, seed=7).sums,
        "estimate_call": lambda: S.estimate(he, sch, S.Call(100.0, 0.99), 0.0, [100.0, 0.4], 64, n, seed=7).sums,
        "ml_level": lambda: np.concatenate(S.multilevel_sample_level(he, S.EulerMaruyama(1 / 16), S.EulerMaruyama(1 / 64), 4, 0.0,
                                                                      [100.0, 0.4], 16, 7000, seed=8)),
        "ml_estimate": lambda: S.ml_estimate(he, S.MultilevelScheme(S.EulerMaruyama(1 / 8), 2, range(0, 5)), S.Call(100.0, 0.99), 0.0,
                                             [100.0, 0.4], 8, [40000, 20000, 9000, 5000, 2500], seed=9).sums,
    }
    w = O.wiener_z(1 / 48, np.random.default_rng(2).standard_normal((5000, 48, 2)))
    runs["simulate_wiener"] = lambda: S.simulate_(np.empty_like(w), he, S.EulerMaruyama(1 / 48), 0.0, [100.0, 0.4], w, math="strict")
    runs["wiener"] = lambda: S.wiener_(np.zeros((6000, 33)), 1 / 32, seed=4)
    sharded = {k: np.array(f(), copy=True) for k, f in runs.items()}
    l0 = S.launch_count()
    single = {k: np.array(_single(S, f), copy=True) for k, f in runs.items()}
    assert S.launch_count() > l0
    for k in runs:
        assert np.array_equal(sharded[k], single[k]), k
    assert np.array_equal(sharded["simulate_wiener"], O.simulate_wiener("Heston", 0, HESTON_P, 0.0, 1 / 48, [100.0, 0.4], w))


def test_sharded_domain_error_and_device_pointers(multi):
    import torch
    S, nd = multi
    bad = S.Heston(0.01, 0.5, 0.02, 2.0, -0.9)         # vol-of-vol far too large: the variance goes negative on many paths
    with pytest.raises(S.DomainError):
        S.sample(bad, S.EulerMaruyama(1 / 8), 0.0, [100.0, 0.02], 8, 40000, seed=1)
    with pytest.raises(S.DomainError):
        S.estimate(bad, S.EulerMaruyama(1 / 8), S.Moments(), 0.0, [100.0, 0.02], 8, 40000, seed=1)
    # device pointers are served by the primary device, unsharded
    out = torch.empty((30000, 2), dtype=torch.float64, device="cuda:0")
    he = S.Heston(*HESTON_P)
    S.sample(he, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.4], 16, 30000, seed=2, out=out)
    assert np.array_equal(out.cpu().numpy(), S.sample(he, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.4], 16, 30000, seed=2))


def test_pageable_and_pinned_host_outputs_agree(gpu):
    """run_slabs: pinned host memory is the copy's destination itself; pageable memory (a Julia Array, a numpy array) is
    filled through the library's pinned ring.  Same bits either way, several slabs (> 256 MB) and one."""
    import torch
    S = gpu
    bs, sch = S.BlackScholes(*BS_P), S.EulerMaruyama(1 / 8)
    for n in (3000, 40_000_000):          # 8 B per path: 320 MB -> two slabs
        pin = torch.empty((n,), dtype=torch.float64, pin_memory=True)
        S.sample(bs, sch, 0.0, 100.0, 8, n, seed=3, out=pin.numpy())
        page = np.empty(n)
        S.sample(bs, sch, 0.0, 100.0, 8, n, seed=3, out=page)
        assert np.array_equal(page, pin.numpy())
    x, _ = S.simulate(bs, sch, 0.0, 100.0, 8, 100_000, seed=3)
    assert np.array_equal(x[:, -1], page[:100_000]) and np.all(x[:, 0] == 100.0)


def test_library_owned_device_buffers(gpu):
    S = gpu
    he, sch = S.Heston(*HESTON_P), S.EulerMaruyama(1 / 32)
    buf = S.DeviceBuffer((2000, 33, 2))
    S.simulate(he, sch, 0.0, [100.0, 0.4], 32, 2000, seed=4, out=buf)
    ref, _ = S.simulate(he, sch, 0.0, [100.0, 0.4], 32, 2000, seed=4)
    assert np.array_equal(buf.to_host(), ref)
    assert np.array_equal(buf.to_host(offset=5 * 33 * 2, count=66), ref[5].reshape(-1))      # one path's slice
    prec, nd_, shape, strides = C.c_int(), C.c_int(), (C.c_int64 * 4)(), (C.c_int64 * 4)()
    S._lib.check(S._lib.lib().sdeb200_buffer_info(buf._h, C.byref(prec), C.byref(nd_), shape, strides))
    assert (prec.value, nd_.value, list(shape)[:3], list(strides)[:3]) == (0, 3, [2000, 33, 2], [33 * 2 * 8, 16, 8])
    # a buffer as the Wiener input AND output of simulate! (in place), then logtransition along it — nothing leaves the GPU
    w = S.DeviceBuffer((500, 65))
    bs = S.BlackScholes(*BS_P)
    S.wiener_(_View(w), 1 / 64, seed=6)
    S.simulate_(_View(w), bs, S.EulerMaruyama(1 / 64), 0.0, 100.0, _View(w))
    x = w.to_host()
    assert np.all(x[:, 0] == 100.0) and np.isfinite(x).all()
    buf.free()
    w.free()


class _View:
    """duck-typed tensor over a DeviceBuffer for the in-place drivers (they only need data_ptr / shape / dtype / contiguity)"""
    def __init__(self, b):
        self.b, self.shape, self.dtype = b, b.shape, np.dtype(b.dtype)

    def data_ptr(self):
        return self.b.ptr

    def is_contiguous(self):
        return True


def test_trajectory_drivers_report_non_finite_paths(gpu, O):
    """simulate / simulate! raise DomainError like the reference's sqrt — after writing the trajectories."""
    S = gpu
    bad = S.Heston(0.01, 0.5, 0.02, 2.0, -0.9)
    out = np.zeros((3000, 17, 2))
    with pytest.raises(S.DomainError):
        S.simulate(bad, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.02], 16, 3000, seed=1, out=out)
    assert np.all(out[:, 0] == [100.0, 0.02]) and not np.isfinite(out[:, -1]).all()      # written anyway
    for layout in ("soa",):
        with pytest.raises(S.DomainError):
            S.simulate(bad, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.02], 16, 3000, seed=1, layout=layout)
    w = O.wiener_z(1 / 16, np.random.default_rng(0).standard_normal((600, 16, 2)) * 3)
    x = np.empty_like(w)
    with pytest.raises(S.DomainError):
        S.simulate_(x, bad, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.02], w)
    good = S.Heston(*HESTON_P)
    S.simulate_(x, good, S.EulerMaruyama(1 / 1024), 0.0, [100.0, 0.4], w * 0.01)        # fine: no error


def test_level_streams_are_namespaced(gpu):
    """level pair i of a multilevel run draws from SDEB200_LEVEL_STREAM(stream, i) = stream + (i << 24): a plain run on
    stream 1 no longer shares its noise with level 1 of a run on stream 0."""
    S = gpu
    he = S.Heston(*HESTON_P)
    ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, range(0, 3))              # κΔt = 0.625: the variance stays positive
    xs = S.sample(he, ml, 0.0, [100.0, 0.4], 16, 2000, seed=3)
    assert len(xs) == 5
    plain1 = S.sample(he, S.EulerMaruyama(1 / 32), 0.0, [100.0, 0.4], 32, 2000, seed=3, stream=1)
    assert not np.array_equal(xs[2], plain1)                                       # level 1's fine path: other noise
    ns = S.sample(he, S.EulerMaruyama(1 / 32), 0.0, [100.0, 0.4], 32, 2000, seed=3, stream=S.level_stream(0, 1))
    assert np.array_equal(xs[2], ns)                                               # ... namely the namespaced stream's
    assert np.array_equal(xs[0], S.sample(he, S.EulerMaruyama(1 / 16), 0.0, [100.0, 0.4], 16, 2000, seed=3))   # pair 0: plain sample
    e = S.ml_estimate(he, ml, S.Component(0), 0.0, [100.0, 0.4], 16, 2000, seed=3)
    assert abs(e.fine_mean[1] - xs[2][:, 0].mean()) <= 1e-9 * 100 and abs(e.fine_mean[2] - xs[4][:, 0].mean()) <= 1e-9 * 100
    with pytest.raises(ValueError):
        S.level_stream(1 << 24, 1)
