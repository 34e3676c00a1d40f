"""GPU tests of the @sde_model -> NVRTC path: functors generated from equation text, compiled at run time,
driven through the same kernels as the built-ins.  The reference's own model / drift / diffusion / simulation
testsets (test/runtests.jl:50-97) are replayed against the device."""
import math

import numpy as np
import pytest

from conftest import BS_P, HESTON_P
from test_dsl import CASES, MODELS

pytestmark = pytest.mark.gpu
EM, MIL, RK = 0, 1, 2


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


_cache = {}


def gen(gpu, name):
    if name not in _cache:
        _cache[name] = gpu.sde_model(name + "_g", MODELS.get(name) or gpu.REGISTRY[name])
    return _cache[name]


def test_reference_drift_testset_on_device(gpu):
    # runtests.jl:61-72 (corrected_drift is outside the hot path)
    S = gpu
    m1, m2 = gen(S, "BlackScholes")(*BS_P), gen(S, "Heston")(*HESTON_P)
    m3 = gen(S, "TestModel")(1, 2, 3, 4, 5, 6)
    assert np.isclose(S.drift(m1, 0.0, 100.0), 100.0 * 0.01)
    assert np.allclose(S.drift(m2, 0.0, [100.0, 0.4]), [0.01 * 100.0, 10.0 * (0.3 - 0.4)])
    assert np.allclose(S.drift(m3, 0.0, [0.1, 0.2, 0.3]), [0.2 * 1, 0.3 * 2, 0.1 * 3])
    assert S.drift(gen(S, "Wiener")(), 0.0, 100.0) == 0.0
    assert S.drift(gen(S, "Deterministic")(), 0.0, 100.0) == 100.0
    td = gen(S, "TimeDependent")(1, 2)
    assert S.drift(td, 0.0, 3.0) != S.drift(td, 1.0, 3.0)
    assert np.isclose(S.drift(td, 4.0, 3.0), 12)
    # the built-in functors answer the same
    assert S.drift(S.BlackScholes(*BS_P), 0.0, 100.0) == S.drift(m1, 0.0, 100.0)
    assert np.array_equal(S.drift(S.Heston(*HESTON_P), 0.0, [100.0, 0.4]), S.drift(m2, 0.0, [100.0, 0.4]))


def test_reference_diffusion_testset_on_device(gpu):
    # runtests.jl:75-82
    S = gpu
    assert np.isclose(S.diffusion(gen(S, "BlackScholes")(*BS_P), 0.0, 100.0), 100.0 * 0.3)
    want = [[100.0 * math.sqrt(0.4), 0.0], [0.4 * 0.1 * math.sqrt(0.4), math.sqrt(1 - 0.4 ** 2) * 0.1 * math.sqrt(0.4)]]
    assert np.allclose(S.diffusion(gen(S, "Heston")(*HESTON_P), 0.0, [100.0, 0.4]), want)
    assert np.array_equal(S.diffusion(S.Heston(*HESTON_P), 0.0, [100.0, 0.4]),
                          S.diffusion(gen(S, "Heston")(*HESTON_P), 0.0, [100.0, 0.4]))
    assert np.allclose(S.diffusion(gen(S, "TestModel")(1, 2, 3, 4, 5, 6), 0.0, [0.1, 0.2, 0.3]),
                       [[4, 5, 0, 0], [0, 1, 1, 0], [0, 0, 6, -1]])
    assert S.diffusion(gen(S, "Wiener")(), 0.0, 100.0) == 1.0
    assert S.diffusion(gen(S, "Deterministic")(), 0.0, 100.0) == 0.0


def test_reference_simulation_testset(gpu):
    # runtests.jl:84-97 (numpy order: Julia (2, 501) == numpy (501, 2))
    S = gpu
    m1, m2 = gen(S, "BlackScholes")(*BS_P), gen(S, "Heston")(*HESTON_P)
    assert S.simulate(m1, S.EulerMaruyama(0.01), 0.0, 100.0, 1000)[0].shape == (1001,)
    assert S.simulate(m1, S.Milstein(0.01), 0.0, 100.0, 1000)[0].shape == (1001,)
    assert S.simulate(m2, S.EulerMaruyama(0.01), 0.0, [100.0, 0.4], 500)[0].shape == (501, 2)
    det = gen(S, "Deterministic")()
    assert np.isclose(S.sample(det, S.EulerMaruyama(1), 0.0, 1.0, 1), 2)
    assert np.allclose(S.simulate(det, S.EulerMaruyama(1.0), 0.0, 1.0, 5)[0], [1, 2, 4, 8, 16, 32])
    assert S.simulate(gen(S, "SpecialCaseModel1")(), S.EulerMaruyama(0.01), 0.0, [100.0], 1000)[0].shape == (1001, 1)
    assert S.simulate(gen(S, "SpecialCaseModel2")(), S.EulerMaruyama(0.01), 0.0, [100.0, 100.0], 1000)[0].shape == (1001, 2)
    assert S.simulate(gen(S, "SpecialCaseModel3")(), S.EulerMaruyama(0.01), 0.0, [100.0] * 3, 1000)[0].shape == (1001, 3)
    # Milstein has no method for M != 1 (runtests.jl:90,92 are commented out for that reason)
    with pytest.raises(S.SDEB200Error):
        S.sample(det, S.Milstein(1.0), 0.0, 1.0, 1)
    with pytest.raises(S.SDEB200Error):
        S.sample(m2, S.RungeKutta(0.01), 0.0, [100.0, 0.4], 10, 10)  # runge_kutta.jl:4: AbstractSDE{D,1} only


@pytest.mark.parametrize("name,params,x,t", CASES)
def test_generated_models_match_oracle_path_by_path(gpu, O, name, params, x, t):
    """Every model: RNG-driven terminal values against the oracle's sample loop on the same normal stream
    (strict math, then fast math), Euler–Maruyama and — where the reference defines it — Milstein."""
    S = gpu
    D, M, _ = O.dims(name)
    m = gen(S, name)(*params)
    nsteps, npaths, dt = 48, 130, 1.0 / 64
    x0 = x[0] if D == 1 else x
    schemes = [EM] + ([MIL, RK] if M == 1 else [])
    for sid in schemes:
        z = O.normals_paths(77, 5, 0, npaths, nsteps, M)
        ref = O.sample_z(name, sid, params, t, dt, x, z)
        sch = (S.EulerMaruyama, S.Milstein, S.RungeKutta)[sid](dt)
        for mth in ("strict", "fast"):
            got = S.sample(m, sch, t, x0, nsteps, npaths, seed=77, stream=5, math=mth).reshape(npaths, D)
            scale = np.maximum(np.abs(ref), 1e-3)
            assert np.max(np.abs(got - ref) / scale) <= 1e-10, (name, sid, mth)


@pytest.mark.parametrize("name,params,x,t", [c for c in CASES if c[0] in
                                             ("Heston", "BlackScholes", "TimeDependent", "CoxIngersollRoss",
                                              "FitzHughNagumo", "Lorenz", "SpecialCaseModel3", "TestModel")])
def test_generated_models_wiener_driven_strict_bitwise(gpu, O, name, params, x, t):
    """simulate!(x, model, scheme, t0, x0, w) in strict math: bit-identical to the oracle for generated code."""
    S = gpu
    D, M, _ = O.dims(name)
    m = gen(S, name)(*params)
    nsteps, npaths, dt = 40, 70, 1.0 / 128
    z = np.random.default_rng(5).standard_normal((npaths, nsteps, M))
    w = O.wiener_z(dt, z)
    for sid in [EM] + ([MIL, RK] if M == 1 else []):
        ref = O.simulate_wiener(name, sid, params, t, dt, x, w)
        out = np.empty_like(ref)
        xs = out[..., 0] if D == 1 else out
        ws = np.ascontiguousarray(w[..., 0] if (M == 1 and D == 1) else w)
        sch = (S.EulerMaruyama, S.Milstein, S.RungeKutta)[sid](dt)
        S.simulate_(xs, m, sch, t, x[0] if D == 1 else x, ws, math="strict")
        assert np.array_equal(out, ref), (name, sid)


def test_generated_heston_equals_builtin(gpu):
    """The hand-written Heston functor and the generated one agree: bit-identical in strict math, <= 1e-12 in
    fast math (different but algebraically equal forms)."""
    S = gpu
    g, b = gen(S, "Heston")(*HESTON_P), S.Heston(*HESTON_P)
    sch = S.EulerMaruyama(1 / 256)
    a1 = S.sample(g, sch, 0.0, [100.0, 0.4], 256, 3000, seed=3, math="strict")
    a2 = S.sample(b, sch, 0.0, [100.0, 0.4], 256, 3000, seed=3, math="strict")
    assert np.array_equal(a1, a2)
    f1 = S.sample(g, sch, 0.0, [100.0, 0.4], 256, 3000, seed=3)
    f2 = S.sample(b, sch, 0.0, [100.0, 0.4], 256, 3000, seed=3)
    assert rel_err(f1, f2) <= 1e-12 and rel_err(f1, a1) <= 1e-12


def test_generated_model_full_paths_and_mlmc(gpu, O):
    S = gpu
    cir = (1.3, 0.09, 0.1)  # far from the r = 0 boundary: no sqrt of a negative rate on either level
    m = gen(S, "CoxIngersollRoss")(*cir)
    sch = S.Milstein(1 / 64)
    x, t = S.simulate(m, sch, 0.0, 0.09, 64, 500, seed=2)
    xs, _ = S.simulate(m, sch, 0.0, 0.09, 64, 500, seed=2, layout="soa")
    assert np.array_equal(x, xs.T) and np.all(x[:, 0] == 0.09)
    assert np.array_equal(x[:, -1], S.sample(m, sch, 0.0, 0.09, 64, 500, seed=2))
    # coupled levels for a generated scalar model
    coarse, fine = S.EulerMaruyama(1 / 16), S.EulerMaruyama(1 / 64)
    xc, xf = S.multilevel_sample_level(m, coarse, fine, 4, 0.0, 0.09, 16, 300, seed=9)
    z = O.normals_paths(9, 0, 0, 300, 64, 1)
    rc, rf = O.multilevel_sample_z("CoxIngersollRoss", EM, cir, 0.0, 1 / 16, 1 / 64, 4, [0.09], z)
    assert rel_err(xc, rc[:, 0]) <= 1e-11 and rel_err(xf, rf[:, 0]) <= 1e-11


def test_coupled_sampler_needs_matching_dimensions(gpu):
    # Δw = zero(typeof(x0)) in multilevel.jl:79 only works when D == M
    S = gpu
    m = gen(S, "TestModel")(1, 2, 3, 4, 5, 6)
    with pytest.raises(S.SDEB200Error) as ei:
        S.multilevel_sample_level(m, S.EulerMaruyama(0.1), S.EulerMaruyama(0.05), 2, 0.0, [0.1, 0.2, 0.3], 4, 10)
    assert ei.value.code == S._lib.E_UNSUPPORTED


@pytest.mark.parametrize("name,params,x,t", [c for c in CASES if c[0] in
                                             ("BlackScholes", "Heston", "Lorenz", "CoxIngersollRoss", "TimeDependent", "FitzHughNagumo")])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_logtransition_matches_oracle(gpu, O, name, params, x, t, precision):
    """logtransition along stored trajectories (schemes.jl:64-76, euler_maruyama.jl:8-13): built-in and generated
    functors, ragged sizes (rows not a multiple of the tile, paths not a multiple of the warp), host and device data."""
    import torch
    S = gpu
    D, M, _ = O.dims(name)
    m = (getattr(S, name) if name in ("BlackScholes", "Heston") else gen(S, name))(*params)
    dt = 1.0 / 256
    for npaths, nsteps in ((1, 1), (37, 5), (100, 45)):
        z = np.random.default_rng(3).standard_normal((npaths, nsteps, M))
        paths = O.simulate_z(name, 0, params, t, dt, x, z)
        data = paths[..., 0] if D == 1 else paths
        data = np.ascontiguousarray(data.astype(np.float32 if precision == "f32" else np.float64))
        ref = O.logtransition(name, params, t, dt, data.reshape(npaths, nsteps + 1, D).astype(np.float64))
        got = S.logtransition(m, S.EulerMaruyama(dt), t + dt * np.arange(nsteps + 1), data)
        assert got.shape == (npaths, nsteps) and got.dtype == data.dtype
        tol = 1e-11 if precision == "f64" else 2e-5
        assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= tol, (name, npaths, nsteps)
        dev = S.logtransition(m, S.EulerMaruyama(dt), t, torch.from_numpy(data).cuda())
        assert np.array_equal(dev.cpu().numpy(), got)
    one = S.logtransition(m, S.EulerMaruyama(dt), t, data[0])
    assert one.shape == (nsteps,) and np.array_equal(one, got[0])


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("npaths,nrows", [(70, 41), (257, 130)])
def test_four_dimensional_model_on_tma_tiles(gpu, precision, npaths, nrows):
    """D = M = 4 (32-byte rows in FP64: a row spans two 16-byte chunks of the swizzled tile; 16-byte rows in FP32): the
    TMA form of simulate!(x, model, scheme, t0, x0, w) equals the plain tile kernel bit for bit, in place and not."""
    import torch
    S = gpu
    cls = S.sde_model("Four_g", "dA = a*A*dt + s*A*dW1\ndB = b*(A - B)*dt + s*dW2\ndC = -c*C*dt + s*B*dW3\ndE = 0.001*A*C*dt + s*dW4")
    assert (cls._D, cls._M) == (4, 4)
    m = cls(0.05, 1.5, 0.7, 0.2)   # fields (a, b, c, s): drift symbols first, then diffusion, in order of appearance
    dt = 1.0 / (nrows - 1)
    tdt = torch.float64 if precision == "f64" else torch.float32
    rng = np.random.default_rng(21)
    w = np.cumsum(np.concatenate([np.zeros((npaths, 1, 4)), rng.standard_normal((npaths, nrows - 1, 4)) * math.sqrt(dt)], axis=1), axis=1)
    res = []
    for tma in (True, False):
        S.set_tma(tma)
        try:
            wd = torch.from_numpy(w).to(tdt).cuda()
            xd = torch.empty_like(wd)
            c0 = S.tma_launch_count()
            S.simulate_(xd, m, S.EulerMaruyama(dt), 0.0, [1.0, 0.5, 0.25, 0.0], wd, math="strict")
            S.simulate_(wd, m, S.EulerMaruyama(dt), 0.0, [1.0, 0.5, 0.25, 0.0], wd, math="strict")   # x === w
            torch.cuda.synchronize()
            res.append((xd.cpu().numpy(), wd.cpu().numpy(), S.tma_launch_count() - c0))
        finally:
            S.set_tma(True)
    assert res[0][2] == 2 and res[1][2] == 0
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][0], res[0][1]) and np.isfinite(res[0][0]).all()
    assert np.all(res[0][0][:, 0] == np.array([1.0, 0.5, 0.25, 0.0], dtype=res[0][0].dtype))


LOGT_TMA_SHAPES = [  # (npaths, nsteps): head rows (odd pitch), leftover paths (npaths % g), partial warps, ragged tails, one tile short
    (257, 512), (1000, 252), (64, 75), (33, 511), (95, 68), (130, 257), (4100, 70), (31, 300),
]


@pytest.mark.parametrize("name,params,x,t", [c for c in CASES if c[0] in
                                             ("BlackScholes", "Heston", "CoxIngersollRoss", "FitzHughNagumo", "Lorenz")])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_logtransition_on_tma_tiles(gpu, O, name, params, x, t, precision):
    """K7t (logtransition with TMA tensor loads, rows of 4 / 8 / 16 bytes; Lorenz' 12- / 24-byte rows stay on K7): bit for bit
    what the plain tile kernel K7 writes, inside guard zones, and the oracle's values (fast logarithm / reciprocal: absolute
    error of the log-determinant <= 4e-16 + 1 ulp, tests/test_device_math.py)."""
    import torch
    S = gpu
    D, M, _ = O.dims(name)
    es = 8 if precision == "f64" else 4
    m = (getattr(S, name) if name in ("BlackScholes", "Heston") else gen(S, name))(*params)
    tdt = torch.float64 if precision == "f64" else torch.float32
    G = 1024
    for npaths, nsteps in LOGT_TMA_SHAPES:
        dt = 1.0 / 256
        z = np.random.default_rng(5).standard_normal((npaths, nsteps, M))
        paths = O.simulate_z(name, 0, params, t, dt, x, z)
        data = np.ascontiguousarray((paths[..., 0] if D == 1 else paths).astype(np.float32 if precision == "f32" else np.float64))
        ref = O.logtransition(name, params, t, dt, data.reshape(npaths, nsteps + 1, D).astype(np.float64))
        nrows = nsteps + 1
        expect_tma = 128 % (D * es) == 0 and nrows >= 2 * (128 // (D * es)) + 4 and npaths >= 32

        def run(tma):
            S.set_tma(tma)
            try:
                fin = torch.full((data.size + 2 * G,), -7.0, dtype=tdt, device="cuda")
                fin[G:G + data.size].copy_(torch.from_numpy(data.reshape(-1)))
                fout = torch.full((npaths * nsteps + 2 * G,), -7.0, dtype=tdt, device="cuda")
                c0 = S.tma_launch_count()
                got = S.logtransition(m, S.EulerMaruyama(dt), t, fin[G:G + data.size].view(*data.shape),
                                      out=fout[G:G + npaths * nsteps].view(npaths, nsteps))
                torch.cuda.synchronize()
                assert bool((fout[:G] == -7.0).all() and (fout[-G:] == -7.0).all()), "guard zone overwritten"
                assert bool((fin[:G] == -7.0).all() and (fin[-G:] == -7.0).all())
                return got.cpu().numpy(), S.tma_launch_count() - c0
            finally:
                S.set_tma(True)

        a, na = run(True)
        b, nb = run(False)
        assert nb == 0 and na == (1 if expect_tma else 0), (name, npaths, nsteps, na)
        assert np.array_equal(a, b, equal_nan=True), (name, npaths, nsteps)
        tol = 1e-11 if precision == "f64" else 2e-5
        assert np.max(np.abs(a - ref) / np.maximum(1.0, np.abs(ref))) <= tol, (name, npaths, nsteps)


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("rng", [1, 2])
def test_reference_order_tma_stores_two_states_one_noise(gpu, precision, rng):
    """D = 2, M = 1 on the TMA store kernel K2u: a Philox batch is several 16-byte chunks (FP32: 4 steps x 8 bytes, FP64
    stream v2: 8 steps x 16 bytes) and, with tiles aligned to 32-byte sectors, its first chunk may sit off the batch grid
    (head rows h_j >= 2), so that a batch straddles two tiles.  Bit for bit what the shared-memory-tile kernel K2b writes."""
    import torch
    S = gpu
    cls = S.sde_model("TwoOne_g", "dX = a*X*dt + s*X*dW1\ndY = b*(X - Y)*dt")
    assert (cls._D, cls._M) == (2, 1)
    m = cls(0.05, 0.3, 1.5)
    tdt = torch.float64 if precision == "f64" else torch.float32
    G = 1024
    for npaths, nrows in [(257, 513), (130, 259), (999, 131), (64, 77), (300, 1026), (97, 70)]:
        nsteps = nrows - 1
        res = []
        for tma in (True, False):
            S.set_tma(tma)
            try:
                n = npaths * nrows * 2
                full = torch.full((n + 2 * G,), -7.0, dtype=tdt, device="cuda")
                v = full[G:G + n].view(npaths, nrows, 2)
                c0 = S.tma_launch_count()
                S.simulate(m, S.EulerMaruyama(1.0 / nsteps), 0.0, [1.0, 0.5], nsteps, npaths, seed=3, precision=precision, out=v, rng=rng)
                torch.cuda.synchronize()
                assert bool((full[:G] == -7.0).all() and (full[-G:] == -7.0).all()), "guard zone overwritten"
                res.append((v.cpu().numpy().copy(), S.tma_launch_count() - c0))
            finally:
                S.set_tma(True)
        assert res[0][1] == 1 and res[1][1] == 0, (npaths, nrows, res[0][1])
        assert np.array_equal(res[0][0], res[1][0]), (npaths, nrows)
        assert np.isfinite(res[0][0]).all() and np.all(res[0][0][:, 0, :] == np.array([1.0, 0.5]))


POLAR_MODELS = [
    # the 3/2 stochastic-volatility model: both noise terms carry sqrt(V), the second one V*sqrt(V)
    ("ThreeHalves", "dS = r*S*dt + sqrt(V)*S*dW1\ndV = κ*V*(θ-V)*dt + σ*V*sqrt(V)*dW2", (0.02, 2.0, 0.3, 0.5), [100.0, 0.3]),
    # a correlated Heston-type model under other names, the root's argument a sum of states
    ("RootOfSum", "dX = a*X*dt + b*sqrt(X+Y)*dW1\ndY = c*(d-Y)*dt + e*sqrt(X+Y)*dW2\ndW1*dW2 = f*dt", (0.05, 0.2, 1.5, 0.4, 0.1, -0.3), [1.0, 0.4]),
]


@pytest.mark.parametrize("name,text,params,x0", POLAR_MODELS)
def test_generated_polar_step_matches_the_independent_twin(gpu, O, name, text, params, x0):
    """Models for which the code generator emits the polar step (one square root of A*w per step, sde_dsl.cpp) against
    oracle/twin.py — an independent parser / evaluator of the same equation text: Wiener-driven strict math bit for bit (no
    polar step there), RNG-driven `sample` on both normal streams path by path (the polar step in FAST math; the plain
    step in STRICT math), every kernel form with the same bits, the fine path of a coupled level included."""
    import twin as T
    S = gpu
    m = S.sde_model(name + "_polar", text)(*params)
    L = S._lib.lib()
    L.sdeb200_model_cuda_source.restype = __import__("ctypes").c_char_p
    assert b"step_polar" in L.sdeb200_model_cuda_source(type(m)._handle)
    tm = T.Model(name, text)
    nsteps, npaths, dt = 48, 40, 1.0 / 96
    z = np.random.default_rng(8).standard_normal((npaths, nsteps, 2))
    w = O.wiener_z(dt, z)
    ref = np.array(T.simulate_wiener(tm, EM, params, 0.0, dt, x0, w.tolist()))
    out = np.empty_like(ref)
    S.simulate_(out, m, S.EulerMaruyama(dt), 0.0, x0, w, math="strict")
    assert np.array_equal(out, ref)
    for rng in (1, 2):
        zz = O.normals_paths(21, 3, 1000, npaths, nsteps, 2, rng=rng)
        ref = np.array(T.sample_z(tm, EM, params, 0.0, dt, x0, zz.tolist()))
        for mth in ("fast", "strict"):
            got = S.sample(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, npaths, seed=21, stream=3, path_offset=1000, rng=rng, math=mth)
            assert rel_err(got, ref) <= 1e-11, (rng, mth)
        term = S.sample(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, 700, seed=21, rng=rng)
        xr, _ = S.simulate(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, 700, seed=21, rng=rng)
        xs, _ = S.simulate(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, 700, seed=21, rng=rng, layout="soa")
        assert np.array_equal(xr[:, -1], term) and np.array_equal(np.moveaxis(xs, -1, 0)[:, -1].reshape(term.shape), term)
        xc, xf = S.multilevel_sample_level(m, S.EulerMaruyama(2 * dt), S.EulerMaruyama(dt), 2, 0.0, x0, nsteps // 2, 700, seed=21, rng=rng)
        assert np.array_equal(xf, term)
    t32 = S.sample(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, 700, seed=21, precision="f32")
    t64 = S.sample(m, S.EulerMaruyama(dt), 0.0, x0, nsteps, 700, seed=21)
    assert np.all(np.isfinite(t32)) and abs(float(t32[:, 0].mean()) / float(t64[:, 0].mean()) - 1) < 0.02   # FP32 stream differs: same law
