"""Normal stream v2 (sdeb200_opts.rng = 2: 96 Philox bits per FP64 pair) on the device, through the C ABI.

The stream is specified independently in oracle/philox_spec.c (pair_f64_v2, long-double libm); every RNG-driven kernel
form must consume it in the reference's draw order (path outer, step inner, M draws left to right,
src/simulation.jl:4-11,42-46), so the kernels are checked path by path against the oracle fed with the specified
normals — the same test the v1 stream passes — plus the kernel-form identities (sample == last row of simulate,
coupled fine path == plain sample, 1-path-per-thread == 4-paths-per-thread, sharding) and moments within 3 SE."""
import numpy as np
import pytest

from conftest import BS_P, HESTON_P

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


MODELS = [  # name, params, x0, scheme ids, nsteps (not multiples of the 8-normal batch on purpose)
    ("BlackScholes", BS_P, 100.0, (0, 1, 2), 37),
    ("Heston", HESTON_P, [100.0, 0.4], (0,), 131),
    ("OrnsteinUhlenbeck", (2.0, 0.5, 0.3), 0.2, (0, 1), 64),
]


def _scheme(S, sid, dt):
    return (S.EulerMaruyama, S.Milstein, S.RungeKutta)[sid](dt)


@pytest.mark.parametrize("name,params,x0,sids,nsteps", MODELS)
def test_sample_v2_matches_oracle_on_the_specified_stream(gpu, O, name, params, x0, sids, nsteps):
    S = gpu
    D, M, _ = O.dims(name)
    dt = 1.0 / nsteps
    m = getattr(S, name)(*params)
    for sid in sids:
        for math, tol in (("strict", 1e-11), ("fast", 1e-11)):
            got = S.sample(m, _scheme(S, sid, dt), 0.0, x0, nsteps, 3000, seed=11, stream=2, rng=2, math=math)
            z = O.normals_paths(11, 2, 0, 48, nsteps, M, rng=2)
            ref = O.sample_z(name, sid, params, 0.0, dt, x0, z)
            assert rel_err(got.reshape(3000, D)[:48], ref) <= tol, (name, sid, math)
            # the last paths of the launch too (chunk tails, global path index as the counter)
            z = O.normals_paths(11, 2, 2990, 10, nsteps, M, rng=2)
            ref = O.sample_z(name, sid, params, 0.0, dt, x0, z)
            assert rel_err(got.reshape(3000, D)[2990:], ref) <= tol


def test_v2_differs_from_v1_and_fp32_is_shared(gpu):
    S = gpu
    m, sch = S.Heston(*HESTON_P), S.EulerMaruyama(1 / 64)
    a = S.sample(m, sch, 0.0, [100.0, 0.4], 64, 512, seed=1, rng=1)
    b = S.sample(m, sch, 0.0, [100.0, 0.4], 64, 512, seed=1, rng=2)
    d = S.sample(m, sch, 0.0, [100.0, 0.4], 64, 512, seed=1)          # default == v1
    assert np.array_equal(a, d) and not np.array_equal(a, b)
    a32 = S.sample(m, sch, 0.0, [100.0, 0.4], 64, 512, seed=1, rng=1, precision="f32")
    b32 = S.sample(m, sch, 0.0, [100.0, 0.4], 64, 512, seed=1, rng=2, precision="f32")
    assert np.array_equal(a32, b32)


def test_v2_kernel_forms_agree(gpu):
    """sample == last row of simulate (both layouts); 1-path-per-thread == 4-paths-per-thread; shards == one launch;
    fine path of the coupled sampler == plain sample; wiener! increments == the stream."""
    S = gpu
    for name, params, x0, nsteps in (("Heston", HESTON_P, [100.0, 0.4], 45), ("BlackScholes", BS_P, 100.0, 70)):
        m, sch = getattr(S, name)(*params), S.EulerMaruyama(1.0 / nsteps)
        n = 1500
        term = S.sample(m, sch, 0.0, x0, nsteps, n, seed=3, rng=2)
        xr, _ = S.simulate(m, sch, 0.0, x0, nsteps, n, seed=3, rng=2)
        xs, _ = S.simulate(m, sch, 0.0, x0, nsteps, n, seed=3, rng=2, layout="soa")
        assert np.array_equal(xr[:, -1], term)
        assert np.array_equal(np.moveaxis(xs, -1, 0)[:, -1].reshape(term.shape), term)
        prev = S.set_sample1_max(0)
        try:
            term4 = S.sample(m, sch, 0.0, x0, nsteps, n, seed=3, rng=2)
        finally:
            S.set_sample1_max(prev)
        assert np.array_equal(term4, term)
        parts = [S.sample(m, sch, 0.0, x0, nsteps, 1024, seed=3, rng=2), S.sample(m, sch, 0.0, x0, nsteps, n - 1024, seed=3, rng=2, path_offset=1024)]
        assert np.array_equal(np.concatenate(parts), term)
        coarse = S.EulerMaruyama(2.0 / nsteps)
        if nsteps % 2 == 0:
            xc, xf = S.multilevel_sample_level(m, coarse, sch, 2, 0.0, x0, nsteps // 2, n, seed=3, rng=2)
            assert np.array_equal(xf, term)
    w = np.zeros((4, 41, 2))
    S.wiener_(w, 1.0, seed=5, stream=1, path_offset=77, rng=2)
    import oracle as O
    zz = np.diff(w[0], axis=0).reshape(-1)
    assert np.max(np.abs(zz - O.normals(5, 1, 77, 0, 80, rng=2))) <= 1e-12


def test_v2_moments_and_estimator(gpu):
    """Heston EM moments within 3 standard errors of the closed forms on stream v2; the fused estimator
    equals the mean of the sampled states and is invariant under sharding."""
    S = gpu
    N, n = 256, 1 << 20
    m, sch = S.Heston(*HESTON_P), S.EulerMaruyama(1.0 / N)
    e = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], N, n, seed=21, rng=2)
    ES = 100.0 * (1 + 0.01 / N) ** N
    EV = 0.3 + (0.4 - 0.3) * (1 - 10.0 / N) ** N
    assert abs(e.mean[0] - ES) <= 3 * e.stderr[0] and abs(e.mean[1] - EV) <= 3 * e.stderr[1]
    x = S.sample(m, sch, 0.0, [100.0, 0.4], N, n, seed=21, rng=2)
    assert abs(e.mean[0] - x[:, 0].mean()) <= 1e-9 * ES
    a = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], N, n // 2, seed=21, rng=2)
    b = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], N, n // 2, seed=21, rng=2, path_offset=n // 2)
    assert abs((a.sums + b.sums)[0, 0] - e.sums[0, 0]) <= 1e-9 * abs(e.sums[0, 0])


def test_v2_statistical_battery(gpu):
    """Moments to 8th order, Kolmogorov–Smirnov, pair correlations and lag correlations of the raw normals of stream v2
    (2.7e8 draws through wiener! with dt = 1)."""
    import torch
    from scipy import stats
    S = gpu
    npaths, nrows = 1 << 18, 513
    w = torch.empty((npaths, nrows, 2), dtype=torch.float64, device="cuda")
    S.wiener_(w, 1.0, seed=99, rng=2)
    z = (w[:, 1:] - w[:, :-1]).reshape(-1)          # 2.7e8 normals (differences of running sums: exact to ~1e-14)
    n = z.numel()
    mom = [float((z ** k).mean()) for k in range(1, 9)]
    exact = [0, 1, 0, 3, 0, 15, 0, 105]
    var_of = [1, 2, 15, 96, 945, 10170, 135135, 2016000]      # Var(z^k) = E z^2k - (E z^k)^2
    for k in range(8):
        assert abs(mom[k] - exact[k]) <= 4.5 * np.sqrt(var_of[k] / n), (k + 1, mom[k])
    zz = (w[:, 1:] - w[:, :-1])
    c01 = float((zz[..., 0] * zz[..., 1]).mean())              # the two normals of one pair
    lag = float((zz[:, 1:, 0] * zz[:, :-1, 0]).mean())         # consecutive pairs of a path
    cross = float((zz[1:, :, 0] * zz[:-1, :, 0]).mean())       # neighbouring paths
    for c in (c01, lag, cross):
        assert abs(c) <= 4.5 / np.sqrt(n / 2)
    sub = z[:: 64][: 2_000_000].cpu().numpy()
    assert stats.kstest(sub, "norm").pvalue > 1e-3
    # tails: P(|z| > 4) and P(|z| > 5)
    for t in (4.0, 5.0):
        p = float((z.abs() > t).double().mean())
        pe = 2 * stats.norm.sf(t)
        assert abs(p - pe) <= 5 * np.sqrt(pe / n)
