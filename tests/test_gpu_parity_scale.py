"""Parity at the sizes SURVEY §8(d) and VERDICT r1 ask for (round 1 compared 257-300 paths and coupled levels of <= 40 fine
steps): 10^4 paths against the oracle, the deep levels 6, 7, 8 of C4's MultilevelScheme(EulerMaruyama(1/16), 2, 0:8)
(multilevel.jl:74-89: 1024 / 2048 / 4096 fine steps through the unrolled substeps == 2 branch of the coupled kernel) on
1000 paths, and substeps = 3 (the general branch, with a Philox batch that straddles coarse steps).  Both normal streams."""
import numpy as np
import pytest

from conftest import BS_P, HESTON_P

pytestmark = pytest.mark.gpu
EM, MIL = 0, 1


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("name,params,x0,sid,nsteps", [("BlackScholes", BS_P, 100.0, MIL, 512), ("Heston", HESTON_P, [100.0, 0.4], EM, 1024)])
def test_ten_thousand_paths_wiener_driven(gpu, O, name, params, x0, sid, nsteps):
    """simulate!(x, model, scheme, t0, x0, w), 10^4 paths at the configurations' step counts: STRICT bit-identical, FAST <= 1e-12"""
    S = gpu
    D, M, _ = O.dims(name)
    dt = 1.0 / nsteps
    w = O.wiener_z(dt, np.random.default_rng(17).standard_normal((10_000, nsteps, M)))
    ref = O.simulate_wiener(name, sid, params, 0.0, dt, x0, w)
    m = getattr(S, name)(*params)
    sch = (S.EulerMaruyama, S.Milstein)[sid](dt)
    for math, exact in (("strict", True), ("fast", False)):
        x = np.empty_like(ref)
        S.simulate_(x[..., 0] if D == 1 else x, m, sch, 0.0, x0, np.ascontiguousarray(w[..., 0] if M == 1 else w), math=math)
        if exact:
            assert np.array_equal(x, ref)
        else:
            assert rel_err(x, ref) <= 1e-12


@pytest.mark.parametrize("rng", [1, 2])
def test_ten_thousand_paths_rng_driven(gpu, O, rng):
    """sample(...) on 10^4 paths, every path against the oracle fed with the specified normals (C1's and C2's models)"""
    S = gpu
    for name, params, x0, sid, nsteps in (("BlackScholes", BS_P, 100.0, EM, 252), ("Heston", HESTON_P, [100.0, 0.4], EM, 256)):
        D, M, _ = O.dims(name)
        dt = 1.0 / nsteps
        z = O.normals_paths(31, 5, 123456, 10_000, nsteps, M, rng=rng)
        ref = O.sample_z(name, sid, params, 0.0, dt, x0, z)
        got = S.sample(getattr(S, name)(*params), S.EulerMaruyama(dt), 0.0, x0, nsteps, 10_000, seed=31, stream=5, path_offset=123456, rng=rng)
        assert rel_err(got.reshape(10_000, D), ref) <= 1e-11, name
        e = S.estimate(getattr(S, name)(*params), S.EulerMaruyama(dt), S.Moments(), 0.0, x0, nsteps, 10_000, seed=31, stream=5,
                       path_offset=123456, rng=rng)
        assert abs(e.mean[0] - ref[:, 0].mean()) <= 1e-11 * abs(ref[:, 0].mean()) * 10


@pytest.mark.parametrize("level", [6, 7, 8])
@pytest.mark.parametrize("rng", [1, 2])
def test_coupled_levels_six_to_eight(gpu, O, level, rng):
    """C4: level pairs (l-1, l) of MultilevelScheme(EulerMaruyama(1/16), 2, 0:8), nsteps = 16 (T = 1): 16*2^(l-1) coarse and
    16*2^l fine steps, 1000 paths, against the oracle's _multilevel_sample on the specified normals; the call payoff of the
    pair reduced on the device equals the one computed from the oracle's terminal states."""
    S = gpu
    params, x0 = HESTON_P, [100.0, 0.4]
    base = S.EulerMaruyama(1 / 16)
    coarse, fine = S.subdivide(base, 2 ** (level - 1)), S.subdivide(base, 2 ** level)
    assert fine.dt == O.level_dt(1 / 16, 2, level) and coarse.dt == O.level_dt(1 / 16, 2, level - 1)
    ncoarse, npaths = 16 * 2 ** (level - 1), 1000
    stream = S.level_stream(0, level)
    z = O.normals_paths(77, stream, 0, npaths, 2 * ncoarse, 2, rng=rng)
    rc, rf = O.multilevel_sample_z("Heston", EM, params, 0.0, coarse.dt, fine.dt, 2, x0, z)
    he = S.Heston(*params)
    for math in ("fast", "strict"):
        xc, xf = S.multilevel_sample_level(he, coarse, fine, 2, 0.0, x0, ncoarse, npaths, seed=77, stream=stream, rng=rng, math=math)
        assert rel_err(xf, rf) <= 1e-11 and rel_err(xc, rc) <= 1e-11, (level, math)
    # the whole multilevel estimator uses exactly these streams: its level-`level` sums equal the oracle's
    ml = S.MultilevelScheme(base, 2, range(0, 9))
    n = [0] * 9
    n[level] = npaths
    est = S.ml_estimate(he, ml, S.Call(100.0, 1.0), 0.0, x0, 16, n, seed=77, rng=rng)
    pf, pc = np.maximum(rf[:, 0] - 100.0, 0.0), np.maximum(rc[:, 0] - 100.0, 0.0)
    assert abs(est.sums[level, 0] - (pf - pc).sum()) <= 1e-9 * max(1.0, abs((pf - pc).sum()))
    assert abs(est.sums[level, 2] - pf.sum()) <= 1e-10 * pf.sum()


@pytest.mark.parametrize("name,params,x0", [("Heston", HESTON_P, [100.0, 0.4]), ("BlackScholes", BS_P, 100.0)])
@pytest.mark.parametrize("rng", [1, 2])
def test_coupled_level_substeps_three(gpu, O, name, params, x0, rng):
    """substeps = 3: the general branch of the coupled kernel (no unrolling; Philox batches of 2 / 4 / 8 steps straddle
    the coarse steps), 1000 paths x 243 fine steps; the fine path is the plain sample on the same stream."""
    S = gpu
    D, M, _ = O.dims(name)
    ncoarse, sub, npaths = 81, 3, 1000
    coarse = S.EulerMaruyama(1.0 / ncoarse)
    fine = S.subdivide(coarse, sub)
    z = O.normals_paths(5, 9, 4096, npaths, ncoarse * sub, M, rng=rng)
    rc, rf = O.multilevel_sample_z(name, EM, params, 0.0, coarse.dt, fine.dt, sub, x0, z)
    m = getattr(S, name)(*params)
    xc, xf = S.multilevel_sample_level(m, coarse, fine, sub, 0.0, x0, ncoarse, npaths, seed=5, stream=9, path_offset=4096, rng=rng)
    assert rel_err(xf.reshape(npaths, D), rf) <= 1e-11 and rel_err(xc.reshape(npaths, D), rc) <= 1e-11
    plain = S.sample(m, fine, 0.0, x0, ncoarse * sub, npaths, seed=5, stream=9, path_offset=4096, rng=rng)
    assert np.array_equal(plain, xf)
