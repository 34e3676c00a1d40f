"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.md §5 / north_star): FP64 trajectories on identical Wiener increments <= 1e-12 relative
(bit-identical in STRICT math), FP32 <= 1e-5 relative; on-device RNG estimates within 3 standard errors.
"""
import math

import numpy as np
import pytest

from conftest import BS_P, HESTON_P

pytestmark = pytest.mark.gpu

EM, MIL, RK = 0, 1, 2


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def wiener_paths(O, npaths, nsteps, dim, dt, seed):
    z = np.random.default_rng(seed).standard_normal((npaths, nsteps, dim))
    return O.wiener_z(dt, z)


CASES = [
    ("BlackScholes", BS_P, 100.0, EM, 252),
    ("BlackScholes", BS_P, 100.0, MIL, 512),
    ("BlackScholes", BS_P, 100.0, RK, 128),
    ("Heston", HESTON_P, [100.0, 0.4], EM, 1024),
]


def _model(gpu, name, params):
    return getattr(gpu, name)(*params)


def _scheme(gpu, sid, dt):
    return (gpu.EulerMaruyama, gpu.Milstein, gpu.RungeKutta)[sid](dt)


@pytest.mark.parametrize("name,params,x0,sid,nsteps", CASES)
def test_wiener_driven_strict_is_bit_identical(gpu, O, name, params, x0, sid, nsteps):
    """simulate!(x, model, scheme, t0, x0, w): STRICT math reproduces the oracle bit for bit."""
    dt = 1.0 / nsteps
    D, M, _ = O.dims(name)
    w = wiener_paths(O, 257, nsteps, M, dt, 1)
    ref = O.simulate_wiener(name, sid, params, 0.0, dt, x0, w)
    x = np.empty_like(ref)
    xs = x[..., 0] if D == 1 else x
    ws = w[..., 0] if M == 1 else w
    gpu.simulate_(xs, _model(gpu, name, params), _scheme(gpu, sid, dt), 0.0, x0, np.ascontiguousarray(ws), math="strict")
    assert np.array_equal(x, ref)


@pytest.mark.parametrize("name,params,x0,sid,nsteps", CASES)
def test_wiener_driven_fast_within_1e12(gpu, O, name, params, x0, sid, nsteps):
    dt = 1.0 / nsteps
    D, M, _ = O.dims(name)
    w = wiener_paths(O, 300, nsteps, M, dt, 2)
    ref = O.simulate_wiener(name, sid, params, 0.0, dt, x0, w)
    x = np.empty_like(ref)
    xs = x[..., 0] if D == 1 else x
    ws = w[..., 0] if M == 1 else w
    gpu.simulate_(xs, _model(gpu, name, params), _scheme(gpu, sid, dt), 0.0, x0, np.ascontiguousarray(ws), math="fast")
    assert rel_err(x, ref) <= 1e-12


@pytest.mark.parametrize("name,params,x0,sid,nsteps", CASES)
def test_wiener_driven_f32_within_1e5(gpu, O, name, params, x0, sid, nsteps):
    """north_star: FP32 trajectories within 1e-5 of the reference arithmetic ON IDENTICAL INCREMENTS: the Wiener path is
    float-valued (drawn with the float oracle's wiener!), so w[i] - w[i-1] is exact in float and in double; every row of
    every path is compared with the Float64 oracle.  (tests/test_gpu_fp32_parity.py pins the FP32 kernels bit for bit
    to a float restatement of the same loops and records the per-row deviation table.)"""
    dt = 1.0 / nsteps
    D, M, _ = O.dims(name)
    z = np.random.default_rng(3).standard_normal((2000, nsteps, M)).astype(np.float32)
    w = O.wiener_z(dt, z, precision="f32")
    ref = O.simulate_wiener(name, sid, params, 0.0, dt, x0, w.astype(np.float64))
    for math in ("strict", "fast"):
        x = np.empty(ref.shape, dtype=np.float32)
        xs = x[..., 0] if D == 1 else x
        ws = w[..., 0] if M == 1 else w
        gpu.simulate_(xs, _model(gpu, name, params), _scheme(gpu, sid, dt), 0.0, x0, np.ascontiguousarray(ws), math=math)
        assert rel_err(x, ref) <= 1e-5, (name, sid, math)


def test_wiener_driven_in_place(gpu, O):
    """x === w is allowed (simulation.jl:120)."""
    name, params, x0, nsteps = "Heston", HESTON_P, [100.0, 0.4], 64
    dt = 1.0 / nsteps
    w = wiener_paths(O, 40, nsteps, 2, dt, 4)
    ref = O.simulate_wiener(name, EM, params, 0.0, dt, x0, w.copy())
    buf = w.copy()
    gpu.simulate_(buf, _model(gpu, name, params), gpu.EulerMaruyama(dt), 0.0, x0, buf, math="strict")
    assert np.array_equal(buf, ref)


TMA_SHAPES = [  # (model, npaths, nrows): ragged tails, head rows (odd pitch), leftover paths (npaths % g), partial warps
    ("BlackScholes", 257, 513), ("BlackScholes", 1000, 253), ("BlackScholes", 64, 40), ("BlackScholes", 33, 512),
    ("BlackScholes", 95, 37), ("Heston", 257, 129), ("Heston", 100, 41), ("Heston", 32, 24),
]


@pytest.mark.parametrize("name,npaths,nrows", TMA_SHAPES)
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("inplace", [False, True])
def test_wiener_driven_tma_tiles(gpu, O, name, npaths, nrows, precision, inplace):
    """The TMA tensor-tile form of simulate!(x, model, scheme, t0, x0, w) (K3t): STRICT FP64 bit-identical to the
    oracle, every precision bit-identical to the plain tile kernel (K3) on the same input, device buffers inside
    guard zones (no stray TMA stores), x === w and x !== w."""
    import torch
    params, x0 = (BS_P, 100.0) if name == "BlackScholes" else (HESTON_P, [100.0, 0.4])
    D = 1 if name == "BlackScholes" else 2
    nsteps = nrows - 1
    dt = 1.0 / nsteps
    w = wiener_paths(O, npaths, nsteps, D, dt, 11)
    ref = O.simulate_wiener(name, EM, params, 0.0, dt, x0, w)
    tdt = torch.float64 if precision == "f64" else torch.float32
    G = 1024
    shape = (npaths, nrows, D) if D > 1 else (npaths, nrows)
    n = int(np.prod(shape))

    def guarded():
        full = torch.full((n + 2 * G,), -7.0, dtype=tdt, device="cuda")
        return full[G:G + n].view(*shape), full

    def run(tma):
        gpu.set_tma(tma)
        try:
            wd, fw = guarded()
            wd.copy_(torch.from_numpy(np.ascontiguousarray(w[..., 0] if D == 1 else w)).to(tdt))
            xd, fx = (wd, fw) if inplace else guarded()
            c0 = gpu.tma_launch_count()
            gpu.simulate_(xd, _model(gpu, name, params), gpu.EulerMaruyama(dt), 0.0, x0, wd, math="strict")
            torch.cuda.synchronize()
            for f in (fw, fx):
                assert bool((f[:G] == -7.0).all() and (f[-G:] == -7.0).all()), "guard zone overwritten"
            return xd.cpu().numpy().reshape(npaths, nrows, D), gpu.tma_launch_count() - c0
        finally:
            gpu.set_tma(True)

    x_tma, n_tma = run(True)
    x_old, n_old = run(False)
    assert n_old == 0 and n_tma == (1 if nrows >= 2 * (128 // (D * (8 if precision == "f64" else 4))) + 4 and npaths >= 32 else 0)
    assert np.array_equal(x_tma, x_old)
    if precision == "f64":
        assert np.array_equal(x_tma, ref)


@pytest.mark.parametrize("npaths,nrows", [(257, 513), (1000, 253), (64, 40), (33, 512), (95, 37), (4097, 65), (32, 300),
                                          (130, 258), (70, 131), (999, 1027)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("rng", [1, 2])
def test_reference_order_trajectories_with_tma_stores(gpu, npaths, nrows, precision, rng):
    """simulate / wiener! / the Exact sampler in the reference's [path][time][D] order: the TMA tensor-store kernel (K2u:
    one sub-path per warp, whole 16-byte chunks per Philox batch, rows of 4 / 8 / 16 bytes, both normal streams) writes bit
    for bit what the shared-memory-tile kernel (K2b) writes, inside guard zones, for ragged shapes (every phase of the
    16-byte grid: nrows mod 4, head rows, partial last tiles, leftover paths, partial warps)."""
    import torch
    S = gpu
    tdt = torch.float64 if precision == "f64" else torch.float32
    G = 1024
    nsteps = nrows - 1
    dt = 1.0 / nsteps

    def guarded(shape):
        n = int(np.prod(shape))
        full = torch.full((n + 2 * G,), -7.0, dtype=tdt, device="cuda")
        return full[G:G + n].view(*shape), full

    def intact(f):
        return bool((f[:G] == -7.0).all() and (f[-G:] == -7.0).all())

    he, bs = S.Heston(*HESTON_P), S.BlackScholes(*BS_P)
    jobs = [
        ("bs-milstein", 1, lambda v: S.simulate(bs, S.Milstein(dt), 0.0, 100.0, nsteps, npaths, seed=5, precision=precision, out=v, path_offset=3, rng=rng)),
        ("heston-em", 2, lambda v: S.simulate(he, S.EulerMaruyama(dt), 0.0, [100.0, 0.4], nsteps, npaths, seed=6, precision=precision, out=v, rng=rng)),
        ("bs-exact", 1, lambda v: S.simulate(bs, S.Exact(dt), 0.0, 100.0, nsteps, npaths, seed=7, precision=precision, out=v, rng=rng)),
        ("wiener1", 1, lambda v: S.wiener_(v, dt, seed=8, rng=rng)),
        ("wiener2", 2, lambda v: S.wiener_(v, dt, seed=9, rng=rng)),
        ("wiener4", 4, lambda v: S.wiener_(v, dt, seed=10, rng=rng)),
    ]
    for tag, D, fn in jobs:
        shape = (npaths, nrows, D) if D > 1 else (npaths, nrows)
        res = []
        for tma in (True, False):
            S.set_tma(tma)
            try:
                v, f = guarded(shape)
                c0 = S.tma_launch_count()
                fn(v)
                torch.cuda.synchronize()
                assert intact(f), f"{tag}: guard zone overwritten (tma={tma})"
                res.append((v.cpu().numpy().copy(), S.tma_launch_count() - c0))
            finally:
                S.set_tma(True)
        rowb = D * (8 if precision == "f64" else 4)
        M = D                                                       # every job here has as many Wiener processes as states
        U = (8 // math.gcd(M, 8)) if (rng == 2 and precision == "f64") else ((2 if precision == "f64" else 4) // math.gcd(M, 2 if precision == "f64" else 4))
        bb = U * rowb
        takes_tma = 16 % rowb == 0 and bb % 16 == 0 and 128 % bb == 0 and nrows >= 2 * (128 // rowb) + 4 and npaths >= 32
        assert res[1][1] == 0 and res[0][1] == (1 if takes_tma else 0), (tag, res[0][1], takes_tma)
        assert np.array_equal(res[0][0], res[1][0]), tag
        assert np.isfinite(res[0][0]).all(), tag


@pytest.mark.parametrize("name,params,x0,sid,nsteps", CASES)
@pytest.mark.parametrize("math", ["fast", "strict"])
def test_sample_matches_oracle_on_the_same_normal_stream(gpu, O, name, params, x0, sid, nsteps, math):
    """RNG-driven terminal values, path by path: the oracle consumes the normals of the documented
    Philox stream (oracle/philox_spec.c) through the reference's sample loop (simulation.jl:39-48)."""
    dt = 1.0 / nsteps
    D, M, _ = O.dims(name)
    npaths, seed, stream, off = 300, 12345, 7, 1000
    z = O.normals_paths(seed, stream, off, npaths, nsteps, M)
    ref = O.sample_z(name, sid, params, 0.0, dt, x0, z)
    got = gpu.sample(_model(gpu, name, params), _scheme(gpu, sid, dt), 0.0, x0, nsteps, npaths, seed=seed,
                     stream=stream, path_offset=off, math=math)
    got = got.reshape(npaths, D)
    assert rel_err(got, ref) <= 1e-11


def test_sample_f32_matches_oracle_on_the_same_normal_stream(gpu, O):
    name, params, x0, nsteps = "Heston", HESTON_P, [100.0, 0.4], 256
    dt = 1.0 / nsteps
    npaths = 200
    z = O.normals_paths(9, 0, 0, npaths, nsteps, 2, "f32")
    ref = O.sample_z(name, EM, params, 0.0, dt, x0, z)
    got = gpu.sample(_model(gpu, name, params), gpu.EulerMaruyama(dt), 0.0, x0, nsteps, npaths, seed=9, precision="f32")
    assert got.dtype == np.float32
    assert rel_err(got, ref) <= 5e-4


@pytest.mark.parametrize("name,params,x0,sid,nsteps", CASES)
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_simulate_layouts_agree_and_end_in_sample(gpu, name, params, x0, sid, nsteps, precision):
    """Full trajectories: reference layout == transposed SoA layout, row 0 == x0, last row == sample()."""
    nsteps = min(nsteps, 200)
    dt = 1.0 / nsteps
    m, sch = _model(gpu, name, params), _scheme(gpu, sid, dt)
    npaths = 1000
    xr, t = gpu.simulate(m, sch, 0.0, x0, nsteps, npaths, seed=5, precision=precision)
    xs, _ = gpu.simulate(m, sch, 0.0, x0, nsteps, npaths, seed=5, precision=precision, layout="soa")
    term = gpu.sample(m, sch, 0.0, x0, nsteps, npaths, seed=5, precision=precision)
    D = m._D
    xr3 = xr.reshape(npaths, nsteps + 1, D)
    xs3 = xs.reshape(nsteps + 1, D, npaths)
    assert np.array_equal(xr3, np.transpose(xs3, (2, 0, 1)))
    assert np.array_equal(xr3[:, 0, :], np.broadcast_to(np.atleast_1d(np.asarray(x0, dtype=xr.dtype)), (npaths, D)))
    assert np.array_equal(xr3[:, -1, :], term.reshape(npaths, D))
    assert t.shape == (nsteps + 1,) and t[0] == 0.0 and abs(t[-1] - 1.0) < 1e-12


def test_simulate_trajectory_matches_oracle_rows(gpu, O):
    name, params, x0, nsteps = "BlackScholes", BS_P, 100.0, 77
    dt = 1.0 / nsteps
    npaths = 150
    z = O.normals_paths(3, 2, 0, npaths, nsteps, 1)
    ref = O.simulate_z(name, MIL, params, 0.0, dt, x0, z)[..., 0]
    x, _ = gpu.simulate(_model(gpu, name, params), gpu.Milstein(dt), 0.0, x0, nsteps, npaths, seed=3, stream=2)
    assert rel_err(x, ref) <= 1e-11


def test_device_resident_outputs(gpu):
    import torch
    m, sch = gpu.Heston(*HESTON_P), gpu.EulerMaruyama(1 / 64)
    a = gpu.sample(m, sch, 0.0, [100.0, 0.4], 64, 5000, seed=1)
    b = gpu.sample(m, sch, 0.0, [100.0, 0.4], 64, 5000, seed=1, device=True)
    assert isinstance(b, torch.Tensor) and b.is_cuda
    assert np.array_equal(a, b.cpu().numpy())


def test_sharded_paths_reproduce_the_single_launch(gpu):
    """path_offset: two shards of one logical run == the run itself (multi-GPU partitioning)."""
    m, sch = gpu.Heston(*HESTON_P), gpu.EulerMaruyama(1 / 32)
    full = gpu.sample(m, sch, 0.0, [100.0, 0.4], 32, 5000, seed=8)
    lo = gpu.sample(m, sch, 0.0, [100.0, 0.4], 32, 2048, seed=8)
    hi = gpu.sample(m, sch, 0.0, [100.0, 0.4], 32, 5000 - 2048, seed=8, path_offset=2048)
    assert np.array_equal(full, np.concatenate([lo, hi]))


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("npaths", [1, 31, 1000, 4097, 70000])
def test_small_launch_kernel_form_is_bit_identical(gpu, precision, npaths):
    """Launches too small to fill the GPU use one path per thread (512-thread CTAs) instead of four: same per-path
    arithmetic, same summation tree per 512-path chunk => terminal states, exact sums and non-finite counts agree
    bit for bit with the four-paths-per-thread form (so sharding over GPUs cannot change a result through it)."""
    import torch
    S = gpu
    he, bs = S.Heston(*HESTON_P), S.BlackScholes(*BS_P)
    tdt = torch.float64 if precision == "f64" else torch.float32
    jobs = [(he, S.EulerMaruyama(1 / 64), [100.0, 0.4], S.Moments(), 2), (bs, S.Milstein(1 / 64), 100.0, S.Call(100.0, 0.99), 1),
            (S.Heston(0.01, 10.0, 0.3, 3.0, 0.4), S.EulerMaruyama(0.25), [100.0, 0.05], S.Moments(), 2)]  # goes non-finite
    for m, sch, x0, pay, D in jobs:
        res = []
        for cap in (1 << 40, 0):
            prev = S.set_sample1_max(cap)
            try:
                x = torch.zeros((npaths, D) if D > 1 else (npaths,), dtype=tdt, device="cuda")
                term = torch.zeros_like(x)
                sums, msg = None, ""
                try:
                    S.sample(m, sch, 0.0, x0, 16, npaths, seed=3, precision=precision, path_offset=5, out=x)
                    e = S.estimate(m, sch, pay, 0.0, x0, 16, npaths, seed=3, precision=precision, path_offset=5, terminal=term)
                    sums = np.asarray(e.sums).copy()
                except S.DomainError as err:
                    msg = str(err)
                    try:
                        S.estimate(m, sch, pay, 0.0, x0, 16, npaths, seed=3, precision=precision, path_offset=5, terminal=term)
                    except S.DomainError as err2:
                        msg += str(err2)
                torch.cuda.synchronize()
                res.append((x.cpu().numpy(), term.cpu().numpy(), sums, msg))
            finally:
                S.set_sample1_max(prev)
        assert np.array_equal(res[0][0], res[1][0], equal_nan=True) and np.array_equal(res[0][1], res[1][1], equal_nan=True)
        assert np.array_equal(res[0][0], res[0][1], equal_nan=True)
        assert res[0][3] == res[1][3]   # same non-finite counts in the messages
        if res[0][2] is not None:
            assert np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize("substeps,ncoarse", [(2, 16), (4, 10)])
def test_coupled_level_matches_oracle(gpu, O, substeps, ncoarse):
    """_multilevel_sample (multilevel.jl:65-89): same increments drive fine and coarse paths."""
    name, params, x0 = "Heston", HESTON_P, [100.0, 0.4]
    dtc = 1.0 / ncoarse
    dtf = O.level_dt(dtc, substeps, 1)
    npaths = 200
    z = O.normals_paths(21, 3, 0, npaths, ncoarse * substeps, 2)
    rc, rf = O.multilevel_sample_z(name, EM, params, 0.0, dtc, dtf, substeps, x0, z)
    coarse = gpu.EulerMaruyama(dtc)
    fine = gpu.subdivide(coarse, substeps)
    assert fine.dt == dtf
    xc, xf = gpu.multilevel_sample_level(gpu.Heston(*params), coarse, fine, substeps, 0.0, x0, ncoarse, npaths,
                                         seed=21, stream=3)
    assert rel_err(xf, rf) <= 1e-11 and rel_err(xc, rc) <= 1e-11
    # the fine path of the pair is the plain sample with the fine scheme on the same stream
    plain = gpu.sample(gpu.Heston(*params), fine, 0.0, x0, ncoarse * substeps, npaths, seed=21, stream=3)
    assert np.array_equal(plain, xf)


def test_reference_multilevel_testset(gpu):
    """test/runtests.jl:99-122 replayed: fine path of _multilevel_simulate == simulate!(fine) on the same
    Wiener path; terminals of _multilevel_sample == last rows of _multilevel_simulate (tolerances of the
    reference: sum of squared norms <= 1e-16 resp. 1e-13)."""
    S = gpu
    substeps = 4
    coarse = S.EulerMaruyama(0.1)
    fine = S.subdivide(coarse, substeps)
    x0 = [100.0, 0.4]
    m2 = S.Heston(*HESTON_P)
    xc, xf, tc, tf = S.multilevel_simulate_level(m2, coarse, fine, substeps, 0.0, x0, 10, 100, seed=0)
    assert xf.shape == (100, 41, 2) and xc.shape == (100, 11, 2)
    assert np.allclose(tf, 0.0 + fine.dt * np.arange(41)) and np.array_equal(tc, tf[::4])
    # wiener! consumes the stream like simulate! does: the same seed gives the same fine trajectory
    w = np.zeros((100, 41, 2))
    S.wiener_(w, fine.dt, seed=0)
    x = np.empty_like(w)
    S.simulate_(x, m2, fine, 0.0, x0, w)
    assert np.sum((xf - x) ** 2) <= 1e-16
    yc, yf = S.multilevel_sample_level(m2, coarse, fine, substeps, 0.0, x0, 10, 100, seed=0)
    assert np.sum((xf[:, -1] - yf) ** 2) <= 1e-12 * 100 ** 2  # Σδw vs differenced running sum: rounding only
    assert np.sum((xc[:, -1] - yc) ** 2) <= 1e-12 * 100 ** 2
    ml = S.MultilevelScheme(coarse, 4, range(0, 5))
    xs, ts = S.simulate(m2, ml, 0.0, x0, 1, 1, seed=0)
    assert len(xs) == 9 and len({(round(t[0], 12), round(t[-1], 12)) for t in ts}) == 1
    smp = S.sample(m2, ml, 0.0, x0, 1, 1, seed=0)
    assert len(smp) == 9
    for a, b in zip(xs, smp):
        assert np.sum(np.abs(a[:, -1] - b)) <= 1e-9


def test_estimate_equals_mean_of_sample(gpu):
    """estimate == mean(payoff.(sample(...))) — and the exact accumulator returns the correctly rounded sum."""
    import math
    m, sch = gpu.Heston(*HESTON_P), gpu.EulerMaruyama(1 / 64)
    n = 70001
    x = gpu.sample(m, sch, 0.0, [100.0, 0.4], 64, n, seed=4)
    e = gpu.estimate(m, sch, gpu.Moments(), 0.0, [100.0, 0.4], 64, n, seed=4)
    assert e.n == n and e.nonfinite == 0
    assert abs(e.sums[0, 0] - math.fsum(x[:, 0])) / e.sums[0, 0] <= 1e-13
    assert abs(e.sums[1, 0] - math.fsum(x[:, 1])) / e.sums[1, 0] <= 1e-13
    assert abs(e.sums[0, 1] - math.fsum(x[:, 0] ** 2)) / e.sums[0, 1] <= 1e-13
    disc = math.exp(-0.01)
    c = gpu.estimate(m, sch, gpu.Call(100.0, disc), 0.0, [100.0, 0.4], 64, n, seed=4)
    pay = disc * np.maximum(x[:, 0] - 100.0, 0.0)
    assert abs(c.mean[0] - pay.mean()) <= 1e-10 and abs(c.stderr[0] - pay.std(ddof=1) / math.sqrt(n)) <= 1e-10


def test_estimate_with_more_chunks_than_ctas(gpu):
    """CTAs walk over chunks grid-stride once the grid is capped: sums still equal the sampled states' sums."""
    import math
    m, sch = gpu.BlackScholes(*BS_P), gpu.EulerMaruyama(1 / 8)
    n = 148 * 64 * 512 * 2 + 12345
    x = gpu.sample(m, sch, 0.0, 100.0, 8, n, seed=6)
    e = gpu.estimate(m, sch, gpu.Moments(), 0.0, 100.0, 8, n, seed=6)
    assert e.n == n and abs(e.sums[0, 0] - math.fsum(x)) / e.sums[0, 0] <= 1e-13
    assert abs(e.sums[0, 1] - math.fsum(x * x)) / e.sums[0, 1] <= 1e-13


def test_estimate_is_invariant_under_sharding(gpu):
    """Bit-identical sums for 1 and 2 shards (host-combined limbs stand in for the NCCL all-reduce)."""
    m, sch = gpu.Heston(*HESTON_P), gpu.EulerMaruyama(1 / 16)
    n = 50000
    full = gpu.estimate(m, sch, gpu.Moments(), 0.0, [100.0, 0.4], 16, n, seed=2)
    lo, hi = gpu.shard_range(n, 0, 2), gpu.shard_range(n, 1, 2)
    a = gpu.estimate(m, sch, gpu.Moments(), 0.0, [100.0, 0.4], 16, lo[1] - lo[0], seed=2, path_offset=lo[0])
    b = gpu.estimate(m, sch, gpu.Moments(), 0.0, [100.0, 0.4], 16, hi[1] - hi[0], seed=2, path_offset=hi[0])
    assert a.n + b.n == n
    # per-shard sums are exact sums of the same per-chunk partials: their exact total rounds to the same double
    from fractions import Fraction
    for q in range(2):
        for k in range(2):
            tot = Fraction(a.sums[q, k]) + Fraction(b.sums[q, k])
            assert abs(float(tot) - full.sums[q, k]) <= abs(full.sums[q, k]) * 2.3e-16


def test_c1_blackscholes_mean_within_3se(gpu):
    """C1: BlackScholes EM 10^6 x 252, closed-form E[S_N] = 100 (1 + r Δt)^N."""
    m, sch = gpu.BlackScholes(*BS_P), gpu.EulerMaruyama(1 / 252)
    e = gpu.estimate(m, sch, gpu.Moments(), 0.0, 100.0, 252, 10 ** 6, seed=0)
    assert abs(e.mean[0] - 101.00499666826876) <= 3 * e.stderr[0]
    assert abs(e.stderr[0] - 0.030992) < 0.001
    second = 100.0 ** 2 * ((1 + 0.01 / 252) ** 2 + 0.3 ** 2 / 252) ** 252
    assert abs(e.sums[0, 1] / e.n - second) / second < 5e-3


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_heston_moments_within_3se(gpu, precision):
    m, sch = gpu.Heston(*HESTON_P), gpu.EulerMaruyama(1 / 1024)
    e = gpu.estimate(m, sch, gpu.Moments(), 0.0, [100.0, 0.4], 1024, 2 * 10 ** 6, seed=1, precision=precision)
    assert abs(e.mean[0] - 101.00501177656254) <= 3 * e.stderr[0] + (2e-3 if precision == "f32" else 0)
    assert abs(e.mean[1] - 0.3000043222543304) <= 3 * e.stderr[1] + (2e-5 if precision == "f32" else 0)
    assert e.nonfinite == 0


def test_milstein_rejected_for_heston(gpu):
    with pytest.raises(gpu.SDEB200Error) as ei:
        gpu.sample(gpu.Heston(*HESTON_P), gpu.Milstein(0.01), 0.0, [100.0, 0.4], 10, 10)
    assert ei.value.code == gpu._lib.E_UNSUPPORTED


def test_negative_variance_raises_domain_error(gpu):
    """sqrt(V < 0): Julia throws DomainError; here the non-finite paths are counted and reported."""
    m = gpu.Heston(0.01, 10.0, 0.3, 3.0, 0.4)  # huge vol-of-vol with a coarse step drives V negative
    with pytest.raises(gpu.DomainError):
        gpu.sample(m, gpu.EulerMaruyama(0.25), 0.0, [100.0, 0.05], 8, 4096, seed=0)


def test_zero_variance_start_is_not_an_error(gpu, O):
    """sqrt(0) = 0 in the reference; the branch-free fast sqrt must agree (V0 = 0, then V > 0)."""
    params, x0, nsteps = HESTON_P, [100.0, 0.0], 32
    dt = 1.0 / nsteps
    z = O.normals_paths(4, 0, 0, 64, nsteps, 2)
    ref = O.sample_z("Heston", EM, params, 0.0, dt, x0, z)
    for mth in ("fast", "strict"):
        got = gpu.sample(gpu.Heston(*params), gpu.EulerMaruyama(dt), 0.0, x0, nsteps, 64, seed=4, math=mth)
        assert np.all(np.isfinite(got)) and rel_err(got, ref) <= 1e-11


def test_mlmc_estimator_call_price(gpu):
    """C4 (reduced): Heston call, levels 0..5; the telescoping sum matches the finest-level plain estimate
    within 3 SE and the semi-analytic price 22.3375 within bias + 3 SE."""
    import math
    S = gpu
    m = S.Heston(*HESTON_P)
    ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, range(0, 6))
    n = [2 ** (19 - l) for l in range(6)]
    pay = S.Call(100.0, math.exp(-0.01))
    est = S.ml_estimate(m, ml, pay, 0.0, [100.0, 0.4], 16, n, seed=11)
    assert np.all(est.nonfinite == 0) and list(est.n) == n
    assert np.all(np.abs(est.level_mean[1:]) < 1.0) and est.level_var[-1] < est.level_var[1]
    assert abs(est.mean - 22.3374853590066) <= 3 * est.stderr + 0.15


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("npaths,nsteps", [(1, 1), (31, 3), (1000, 37), (4097, 17)])
def test_no_out_of_bounds_writes(gpu, precision, npaths, nsteps):
    """compute-sanitizer is closed on this pool: every device output is carved out of a larger buffer whose guard
    zones (sentinel values on both sides) must be untouched after the kernels ran, for ragged sizes."""
    import torch
    S = gpu
    tdt = torch.float64 if precision == "f64" else torch.float32
    G = 4096

    def guarded(shape):
        n = int(np.prod(shape))
        full = torch.full((n + 2 * G,), -12345.0, dtype=tdt, device="cuda")
        return full[G:G + n].view(*shape), full

    def intact(full):
        return bool((full[:G] == -12345.0).all() and (full[-G:] == -12345.0).all())

    he, bs = S.Heston(*HESTON_P), S.BlackScholes(*BS_P)
    for m, x0, D, sch in ((he, [100.0, 0.4], 2, S.EulerMaruyama(1 / 64)), (bs, 100.0, 1, S.Milstein(1 / 64))):
        shp = (npaths, D) if D > 1 else (npaths,)
        v, f = guarded(shp)
        S.sample(m, sch, 0.0, x0, nsteps, npaths, seed=1, precision=precision, out=v)
        assert intact(f) and bool(torch.isfinite(v).all())
        shp = (npaths, nsteps + 1, D) if D > 1 else (npaths, nsteps + 1)
        v, f = guarded(shp)
        S.simulate(m, sch, 0.0, x0, nsteps, npaths, seed=1, precision=precision, out=v)
        assert intact(f) and bool(torch.isfinite(v).all())
        w, fw = guarded(shp)
        S.wiener_(w, 1 / 64, seed=2)
        assert intact(fw) and bool((w[:, 0] == 0).all())
        S.simulate_(w, m, sch, 0.0, x0, w)
        assert intact(fw) and bool(torch.isfinite(w).all())
        shp = (nsteps + 1, D, npaths) if D > 1 else (nsteps + 1, npaths)
        v, f = guarded(shp)
        S.simulate(m, sch, 0.0, x0, nsteps, npaths, seed=1, precision=precision, out=v, layout="soa")
        assert intact(f) and bool(torch.isfinite(v).all())
        v, f = guarded((npaths, D) if D > 1 else (npaths,))
        S.estimate(m, sch, S.Moments(), 0.0, x0, nsteps, npaths, seed=1, precision=precision, terminal=v)
        assert intact(f) and bool(torch.isfinite(v).all())


def test_adaptive_mlmc_driver(gpu):
    """The caller one level above the hot path: Giles' adaptive allocation on the per-level sums.
    Heston call, eps = 0.05: the estimate lands within 3 eps of the semi-analytic price, the finer levels get
    geometrically fewer paths, and extending a level reproduces the sums of a single run with the final counts."""
    import math
    S = gpu
    m = S.Heston(*HESTON_P)
    pay = S.Call(100.0, math.exp(-0.01))
    r = S.mlmc_adaptive(m, S.EulerMaruyama(1 / 16), 2, pay, 0.0, [100.0, 0.4], 16, 0.05, seed=3)
    assert abs(r.mean - 22.3374853590066) <= 3 * 0.05
    assert r.stderr <= 0.05 and r.bias_estimate <= 0.05
    assert np.all(np.diff(r.n[1:]) <= 0) and r.n[0] > r.n[-1]
    # the same per-level counts in one shot give the same exact sums
    ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, r.levels)
    one = S.ml_estimate(m, ml, pay, 0.0, [100.0, 0.4], 16, list(r.n), seed=3)
    assert np.allclose(one.level_mean, r.level_mean, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name,params,x0", [("BlackScholes", BS_P, 100.0), ("OrnsteinUhlenbeck", (2.0, 0.7, 0.3), 1.1)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_exact_samplers_match_oracle(gpu, O, name, params, x0, precision):
    """Exact scheme (simulation.jl:50-57,78-89): terminal values and trajectories against the oracle on the same
    N(0,1) stream; device exp/log differ from libm by an ulp or two per step."""
    S = gpu
    m = getattr(S, name)(*params)
    nsteps, npaths, dt = 12, 500, 1.0 / 12
    z = O.normals_paths(31, 2, 0, npaths, nsteps, 1, precision)
    ref = O.simulate_z(name, 3, params, 0.0, dt, x0, z)[..., 0]
    x, t = S.simulate(m, S.Exact(dt), 0.0, x0, nsteps, npaths, seed=31, stream=2, precision=precision)
    xs, _ = S.simulate(m, S.Exact(dt), 0.0, x0, nsteps, npaths, seed=31, stream=2, precision=precision, layout="soa")
    term = S.sample(m, S.Exact(dt), 0.0, x0, nsteps, npaths, seed=31, stream=2, precision=precision)
    tol = 1e-11 if precision == "f64" else 2e-4
    assert rel_err(x, ref) <= tol
    assert np.array_equal(x, xs.T) and np.array_equal(x[:, -1], term)
    for mth in ("strict",):
        ts = S.sample(m, S.Exact(dt), 0.0, x0, nsteps, npaths, seed=31, stream=2, precision=precision, math=mth)
        assert rel_err(ts, ref[:, -1]) <= tol


def test_exact_blackscholes_is_bias_free(gpu):
    """One exact step to T = 1: the call price matches Black–Scholes (12.368267) within 3 SE —
    the cross-check for the discretised schemes."""
    import math
    S = gpu
    m = S.BlackScholes(*BS_P)
    e = S.estimate(m, S.Exact(1.0), S.Call(100.0, math.exp(-0.01)), 0.0, 100.0, 1, 4 * 10 ** 6, seed=5)
    assert abs(e.mean[0] - 12.368267463784079) <= 3 * e.stderr[0]
    mom = S.estimate(m, S.Exact(0.25), S.Moments(), 0.0, 100.0, 4, 4 * 10 ** 6, seed=6)
    assert abs(mom.mean[0] - 100 * math.exp(0.01)) <= 3 * mom.stderr[0]


def test_exact_only_where_the_reference_defines_it(gpu):
    S = gpu
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(S.Heston(*HESTON_P), S.Exact(0.1), 0.0, [100.0, 0.4], 4, 8)
    assert ei.value.code == S._lib.E_UNSUPPORTED
    gen = S.sde_model("BSgen", S.REGISTRY["BlackScholes"])(*BS_P)
    with pytest.raises(S.SDEB200Error):
        S.sample(gen, S.Exact(0.1), 0.0, 100.0, 4, 8)
    ml = S.MultilevelScheme(S.Exact(0.1), 2, range(0, 3))
    with pytest.raises(S.SDEB200Error):   # the coupled sampler needs a step(model, scheme, ...) method
        S.sample(S.BlackScholes(*BS_P), ml, 0.0, 100.0, 2, 4)


def test_large_host_buffers_go_through_the_slab_pipelines(gpu):
    """Host arrays larger than one staging slab (256 MB / 128 MB): the double-buffered device->host pipeline of
    sample/simulate, the strided 2-D copies of the SoA layout and the host path of simulate! must reproduce the
    device-resident results bit for bit."""
    import torch
    S = gpu
    he, bs = S.Heston(*HESTON_P), S.BlackScholes(*BS_P)
    n = 18_000_000                                    # 288 MB of (S, V) terminals: two slabs
    sch = S.EulerMaruyama(1 / 256)
    host = S.sample(he, sch, 0.0, [100.0, 0.4], 4, n, seed=3)
    dev = S.sample(he, sch, 0.0, [100.0, 0.4], 4, n, seed=3, device=True)
    assert np.array_equal(host, dev.cpu().numpy())
    del dev
    n2, steps = 300_000, 128                          # 310 MB of trajectories
    for lay in ("reference", "soa"):
        xh, _ = S.simulate(bs, S.Milstein(1 / steps), 0.0, 100.0, steps, n2, seed=4, layout=lay)
        xd, _ = S.simulate(bs, S.Milstein(1 / steps), 0.0, 100.0, steps, n2, seed=4, layout=lay, device=True)
        assert np.array_equal(xh, xd.cpu().numpy()), lay
        del xd
    w = xh if xh.shape[0] == n2 else np.ascontiguousarray(xh.T)   # any (n2, steps+1) array will do as a Wiener path
    wd = torch.from_numpy(w).cuda()
    S.simulate_(wd, bs, S.EulerMaruyama(1 / steps), 0.0, 100.0, wd)
    wh = w.copy()
    S.simulate_(wh, bs, S.EulerMaruyama(1 / steps), 0.0, 100.0, wh)  # host in place: 2 slabs of 128 MB
    assert np.array_equal(wh, wd.cpu().numpy())
    lt_h = S.logtransition(bs, S.EulerMaruyama(1 / steps), 0.0, wh)
    lt_d = S.logtransition(bs, S.EulerMaruyama(1 / steps), 0.0, wd)
    assert np.array_equal(lt_h, lt_d.cpu().numpy())


def test_empty_and_degenerate_sizes(gpu):
    """nsteps = 0 returns the start point (the reference's loops simply do not run), npaths = 0 returns empty
    arrays and an empty estimate, a single path works through every driver."""
    S = gpu
    he, bs = S.Heston(*HESTON_P), S.BlackScholes(*BS_P)
    sch = S.EulerMaruyama(0.01)
    x = S.sample(he, sch, 0.0, [100.0, 0.4], 0, 5)
    assert np.array_equal(x, np.tile([100.0, 0.4], (5, 1)))
    xs, t = S.simulate(bs, sch, 0.0, 100.0, 0, 3)
    assert xs.shape == (3, 1) and np.all(xs == 100.0) and t.shape == (1,)
    assert S.sample(he, sch, 0.0, [100.0, 0.4], 7, 0).shape == (0, 2)
    assert S.simulate(bs, sch, 0.0, 100.0, 7, 0)[0].shape == (0, 8)
    e = S.estimate(he, sch, S.Moments(), 0.0, [100.0, 0.4], 7, 0)
    assert e.n == 0 and np.all(e.sums == 0)
    one = S.sample(bs, S.Milstein(0.01), 0.0, 100.0, 10)
    assert isinstance(one, float) and np.isfinite(one)
    assert S.simulate(he, sch, 0.0, [100.0, 0.4], 10)[0].shape == (11, 2)
    assert S.logtransition(bs, sch, 0.0, np.array([100.0])).shape == (0,)
