"""The CPU oracle against every numeric known-answer the reference's own tests hold for the hot path
(test/runtests.jl), plus RNG-independent closed forms.  Runs without a GPU."""
import math

import numpy as np

from conftest import BS_P, HESTON_P

EM, MIL = 0, 1


def test_model_dims_and_parameter_counts(O):
    # runtests.jl:51-58
    assert O.dims("BlackScholes") == (1, 1, 2)
    assert O.dims("Heston") == (2, 2, 5)
    assert O.dims("TestModel") == (3, 4, 6)
    assert O.dims("Wiener")[:2] == (1, 1)
    assert O.dims("Deterministic")[:2] == (1, 0)


def test_drift_known_answers(O):
    # runtests.jl:61-71
    assert np.isclose(O.drift("BlackScholes", BS_P, 0.0, 100.0)[0], 100.0 * 0.01)
    assert np.allclose(O.drift("Heston", HESTON_P, 0.0, [100.0, 0.4]), [0.01 * 100.0, 10.0 * (0.3 - 0.4)])
    assert np.allclose(O.drift("TestModel", (1, 2, 3, 4, 5, 6), 0.0, [0.1, 0.2, 0.3]), [0.2 * 1, 0.3 * 2, 0.1 * 3])
    assert O.drift("Wiener", (), 0.0, 100.0)[0] == 0.0
    assert O.drift("Deterministic", (), 0.0, 100.0)[0] == 100.0
    assert O.drift("TimeDependent", (1, 2), 0.0, 3.0)[0] != O.drift("TimeDependent", (1, 2), 1.0, 3.0)[0]
    assert np.isclose(O.drift("TimeDependent", (1, 2), 4.0, 3.0)[0], 12)


def test_diffusion_known_answers(O):
    # runtests.jl:75-82 — the Heston row pins the Cholesky factor folded into the diffusion
    assert np.isclose(O.diffusion("BlackScholes", BS_P, 0.0, 100.0)[0, 0], 100.0 * 0.3)
    want = np.array([[100.0 * math.sqrt(0.4), 0.0],
                     [0.4 * 0.1 * math.sqrt(0.4), math.sqrt(1 - 0.4 ** 2) * 0.1 * math.sqrt(0.4)]])
    assert np.allclose(O.diffusion("Heston", HESTON_P, 0.0, [100.0, 0.4]), want)
    assert np.allclose(O.diffusion("TestModel", (1, 2, 3, 4, 5, 6), 0.0, [0.1, 0.2, 0.3]),
                       [[4, 5, 0, 0], [0, 1, 1, 0], [0, 0, 6, -1]])
    assert O.diffusion("Wiener", (), 0.0, 100.0)[0, 0] == 1.0
    assert O.diffusion("Deterministic", (), 0.0, 100.0)[0, 0] == 0.0


def test_noise_free_trajectories(O):
    # runtests.jl:89,91 — the only numeric trajectory KATs of the reference
    assert np.isclose(O.sample_z("Deterministic", EM, (), 0.0, 1.0, 1.0, np.zeros((1, 1, 1)))[0, 0], 2)
    x = O.simulate_z("Deterministic", EM, (), 0.0, 1.0, 1.0, np.zeros((1, 5, 1)))
    assert np.allclose(x.ravel(), [1, 2, 4, 8, 16, 32])


def test_output_shapes(O):
    # runtests.jl:86-88,94-96 in numpy order
    z = np.zeros((1, 1000, 1))
    assert O.simulate_z("BlackScholes", EM, BS_P, 0.0, 0.01, 100.0, z).shape == (1, 1001, 1)
    assert O.simulate_z("BlackScholes", MIL, BS_P, 0.0, 0.01, 100.0, z).shape == (1, 1001, 1)
    assert O.simulate_z("Heston", EM, HESTON_P, 0.0, 0.01, [100.0, 0.4], np.zeros((1, 500, 2))).shape == (1, 501, 2)
    assert O.simulate_z("SpecialCaseModel1", EM, (), 0.0, 0.01, [100.0], np.zeros((1, 1000, 2))).shape == (1, 1001, 1)
    assert O.simulate_z("SpecialCaseModel2", EM, (), 0.0, 0.01, [100.0] * 2, np.zeros((1, 1000, 1))).shape == (1, 1001, 2)
    assert O.simulate_z("SpecialCaseModel3", EM, (), 0.0, 0.01, [100.0] * 3, np.zeros((1, 1000, 2))).shape == (1, 1001, 3)


def test_multilevel_coupling_identities(O):
    """runtests.jl:99-113 with a shared normal array standing in for seed!(0): (i) the fine path of
    _multilevel_simulate == simulate with the fine scheme, (ii) the terminals of _multilevel_sample == the
    last rows of _multilevel_simulate (tolerance of the reference: Σ‖·‖² <= 1e-16 is for its rounding level;
    values here are O(100) so we allow the same relative level)."""
    substeps, nsteps, npaths = 4, 10, 100
    dtc = 0.1
    dtf = O.level_dt(dtc, substeps, 1)
    assert dtf == 0.1 / 4
    x0 = [100.0, 0.4]
    z = np.random.default_rng(0).standard_normal((npaths, nsteps * substeps, 2))
    xc, xf = O.multilevel_simulate_z("Heston", EM, HESTON_P, 0.0, dtc, dtf, substeps, x0, z)
    x = O.simulate_wiener("Heston", EM, HESTON_P, 0.0, dtf, x0, O.wiener_z(dtf, z))
    assert np.sum((xf - x) ** 2) <= 1e-16
    yc, yf = O.multilevel_sample_z("Heston", EM, HESTON_P, 0.0, dtc, dtf, substeps, x0, z)
    assert np.sum((xf[:, -1] - yf) ** 2) <= 1e-16 * 100 ** 2
    assert np.sum((xc[:, -1] - yc) ** 2) <= 1e-16 * 100 ** 2
    # the plain simulate! driven by the same normals differs from the Wiener-differenced one by rounding only
    x2 = O.simulate_z("Heston", EM, HESTON_P, 0.0, dtf, x0, z)
    assert np.max(np.abs(x2 - xf)) <= 1e-10


def test_multilevel_scheme_levels_span_the_same_interval(O):
    # runtests.jl:115-118: MultilevelScheme(EM(0.1), 4, 0:4), nsteps = 1
    ends = {round(1 * 4 ** l * O.level_dt(0.1, 4, l), 14) for l in range(5)}
    assert ends == {0.1}
    assert O.level_dt(0.1, 4, 3) == 0.1 / 64  # one division by the integer power (schemes.jl:53-54)


def test_npaths_allocator(O):
    # multilevel.jl:12-24, worked by hand for substeps=4, levels 0:4, budget 1000
    assert list(O.npaths(4, 0, 4, 1000)) == [276, 69, 16, 3, 0]
    n = O.npaths(2, 0, 8, 2 ** 20)
    c = O.cost(2, 0, 8, n)
    assert c.sum() <= 2 ** 20 and np.all(np.diff(n) <= 0)
    assert list(O.cost(4, 0, 4, np.array([276, 69, 16, 3, 0]))) == [276, 276, 256, 192, 0]


def test_milstein_only_for_one_wiener_process(O):
    # milstein.jl:4: step(::AbstractSDE{D,1}, ::Milstein, ...)
    import pytest
    with pytest.raises(RuntimeError):
        O.sample_z("Heston", MIL, HESTON_P, 0.0, 0.01, [100.0, 0.4], np.zeros((1, 1, 2)))


def test_step_formulas_by_hand(O):
    """euler_maruyama.jl:1-6 / milstein.jl:4-10 against the formulas written out."""
    r, s = BS_P
    S, dt, dw = 100.0, 1 / 252, 0.037
    em = O.sample_z("BlackScholes", EM, BS_P, 0.0, dt, S, np.array([[[dw / math.sqrt(dt)]]]))[0, 0]
    assert np.isclose(em, (S + (r * S) * dt) + (s * S) * dw, rtol=0, atol=1e-12)
    mil = O.sample_z("BlackScholes", MIL, BS_P, 0.0, dt, S, np.array([[[dw / math.sqrt(dt)]]]))[0, 0]
    assert np.isclose(mil, ((S + (r * S) * dt) + (s * S) * dw) + ((s * (s * S)) * (dw * dw - dt)) / 2, rtol=0, atol=1e-12)
    # runge_kutta.jl:7-9
    rk = O.sample_z("BlackScholes", 2, BS_P, 0.0, dt, S, np.array([[[dw / math.sqrt(dt)]]]))[0, 0]
    sq = math.sqrt(dt)
    xh = (S + (r * S) * dt) + (s * S) * sq
    assert np.isclose(rk, ((S + (r * S) * dt) + (s * S) * dw) + ((s * xh - s * S) * (dw * dw - dt)) / (2 * sq), rtol=0, atol=1e-12)
    # for a diffusion linear in the state the derivative-free scheme equals Milstein up to rounding
    assert abs(rk - mil) < 1e-3 and abs(rk - mil) > 0
    rr, k, th, sg, rho = HESTON_P
    S, V, d1, d2 = 100.0, 0.4, 0.011, -0.02
    sd = math.sqrt(dt)
    out = O.sample_z("Heston", EM, HESTON_P, 0.0, dt, [S, V], np.array([[[d1 / sd, d2 / sd]]]))[0]
    sv = math.sqrt(V)
    assert np.isclose(out[0], (S + (rr * S) * dt) + (sv * S) * d1, rtol=1e-15)
    assert np.isclose(out[1], (V + (k * (th - V)) * dt) + (((sg * sv) * rho) * d1 + ((sg * sv) * math.sqrt(1 - rho * rho)) * d2), rtol=1e-15)


def test_closed_form_moments_of_the_schemes(O):
    """E[S_N] = S0 (1 + rΔt)^N for EM and Milstein, E[V_N] = θ + (V0-θ)(1-κΔt)^N, with the
    baseline's own generator — a statistical check of the CPU baseline path."""
    n, N = 200000, 64
    sums, _, _ = O.sample_rng("Heston", EM, HESTON_P, 0.0, 1 / N, [100.0, 0.4], N, n, seed=3)
    mS, mV = sums[0] / n, sums[2] / n
    sdS = math.sqrt(sums[1] / n - mS * mS)
    assert abs(mS - 100 * (1 + 0.01 / N) ** N) <= 4 * sdS / math.sqrt(n)
    assert abs(mV - (0.3 + 0.1 * (1 - 10 / N) ** N)) <= 4 * math.sqrt(sums[3] / n - mV * mV) / math.sqrt(n)


def test_philox_known_answers(O):
    # Random123 kat_vectors: philox4x32-10
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normal_stream_statistics(O):
    for prec in ("f64", "f32"):
        z = O.normals(1, 0, 3, 0, 400000, prec)
        n = len(z)
        assert abs(z.mean()) < 4 / math.sqrt(n)
        assert abs(z.var() - 1) < 4 * math.sqrt(2 / n)
        assert abs((z ** 3).mean()) < 4 * math.sqrt(15 / n)
        assert abs((z ** 4).mean() - 3) < 4 * math.sqrt(96 / n)
        # pairs (cos, sin branches) are uncorrelated
        assert abs(np.mean(z[0::2] * z[1::2])) < 4 / math.sqrt(n / 2)


def test_exact_samplers(O):
    """sample(model, ::Exact, t0, x0) for BlackScholes (black_scholes.jl:3-13) and OrnsteinUhlenbeck
    (ornstein_uhlenbeck.jl:3-13): formulas by hand, and the bias-free moments of the exact transition."""
    r, s = BS_P
    dt, z = 0.5, 0.3
    got = O.sample_z("BlackScholes", 3, BS_P, 0.0, dt, 100.0, np.array([[[z]]]))[0, 0]
    assert got == math.exp(z * (math.sqrt(dt) * s) + (math.log(100.0) + (r - 0.5 * (s * s)) * dt))
    th, mu, sg = 2.0, 0.7, 0.3
    got = O.sample_z("OrnsteinUhlenbeck", 3, (th, mu, sg), 0.0, dt, 1.1, np.array([[[z]]]))[0, 0]
    m = 1.1 * math.exp(-th * dt) + mu * (1.0 - math.exp(-th * dt))
    d = math.sqrt(1.0 - math.exp(-2 * th * dt)) * sg / math.sqrt(2 * th)
    assert got == m + d * z
    n = 400000
    zz = np.random.default_rng(7).standard_normal((n, 4, 1))
    x = O.sample_z("BlackScholes", 3, BS_P, 0.0, 0.25, 100.0, zz)[:, 0]
    assert abs(x.mean() - 100.0 * math.exp(0.01)) <= 4 * x.std() / math.sqrt(n)          # no discretisation bias
    y = O.sample_z("OrnsteinUhlenbeck", 3, (th, mu, sg), 0.0, 0.25, 1.1, zz)[:, 0]
    assert abs(y.mean() - (1.1 * math.exp(-th) + mu * (1 - math.exp(-th)))) <= 4 * y.std() / math.sqrt(n)
    assert abs(y.var() - sg ** 2 / (2 * th) * (1 - math.exp(-2 * th))) <= 0.01 * y.var()
    import pytest
    with pytest.raises(RuntimeError):   # no Exact method for Heston in the reference
        O.sample_z("Heston", 3, HESTON_P, 0.0, 0.1, [100.0, 0.4], np.zeros((1, 1, 2)))


def test_logtransition_against_scipy(O):
    """logtransition(model, EulerMaruyama, times, data) (schemes.jl:64-76): the oracle against an independent
    evaluation of the same Gaussian log-density (scipy) for D = 1, 2, 3."""
    from scipy.stats import multivariate_normal
    LP = (10.0, 28.0, 8 / 3, 0.1, 0.2, 0.3)
    for name, p, x0, dt, M in (("BlackScholes", BS_P, [100.0], 0.01, 1), ("Heston", HESTON_P, [100.0, 0.4], 0.01, 2),
                               ("Lorenz", LP, [1.0, 2.0, 3.0], 0.001, 3)):
        z = np.random.default_rng(0).standard_normal((3, 7, M))
        x = O.simulate_z(name, 0, p, 0.5, dt, x0, z)
        lt = O.logtransition(name, p, 0.5, dt, x)
        assert lt.shape == (3, 7)
        for k in range(3):
            for i in range(1, 8):
                xa, xb = x[k, i - 1], x[k, i]
                mu = xa + O.drift(name, p, 0.5 + (i - 1) * dt, xa) * dt
                sg = O.diffusion(name, p, 0.5 + (i - 1) * dt, xa)
                want = multivariate_normal(mu, dt * sg @ sg.T).logpdf(xb)
                assert abs(lt[k, i - 1] - want) <= 1e-9 * max(1.0, abs(want))
