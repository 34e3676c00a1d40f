"""The C oracle (oracle/sde_oracle.c) against an independent second restatement (oracle/twin.py): pure Python, models
parsed from the `@sde_model` text and evaluated by substitution as src/models/codegen.jl does, Milstein's derivative by
forward-mode dual numbers as src/schemes/milstein.jl:1-2 does through ForwardDiff.  Everything is compared BIT FOR BIT
(both sides are IEEE double without FMA contraction).  This narrows the round-1 parity cap "one restatement, written by
the author of the kernels, is the sole authority for Milstein / RungeKutta": the hand-coded ∂σ of the C oracle now
have to agree with automatic differentiation of the model text.  CPU only."""
import numpy as np
import pytest

import twin as T
from test_dsl import CASES, MODELS

REGISTRY = {
    "BlackScholes": "dS = r*S*dt + σ*S*dW",
    "Heston": "dS = r*S*dt + sqrt(V)*S*dW1\ndV = κ*(θ-V)*dt + σ*sqrt(V)*dW2\ndW1*dW2 = ρ*dt",
    "CoxIngersollRoss": "dr = κ*(θ-r)*dt + σ*sqrt(r)*dW",
    "OrnsteinUhlenbeck": "dx = θ*(μ-x)*dt + σ*dW",
    "FitzHughNagumo": "dX = ɛ*(X-X^3-Y+s)*dt + σ1*dW1\ndY = (γ*X-Y+β)*dt + σ2*dW2",
    "Lorenz": "dx = σ*(y-x)*dt + σ1*dw1\ndy = (x*(ρ-z)-y)*dt + σ2*dw2\ndz = (x*y-β*z)*dt + σ3*dw3",
}


def twin_model(name):
    return T.Model(name, MODELS.get(name) or REGISTRY[name])


def test_registry_text_is_the_reference_text():
    """the equation strings above are the reference's (src/models/black_scholes.jl:1, models.jl:39-54, ...) and the
    product's registry carries the same text"""
    import re
    import sys
    src = open(__file__.replace("tests/test_twin.py", "sdemodels.jl_b200/models.py"), encoding="utf-8").read()
    for name, eq in REGISTRY.items():
        assert eq.replace("\n", "\\n") in src, name


def test_parameter_order_and_dims_follow_codegen():
    # test/runtests.jl:51-58 and the registry models (order of appearance, diffusion column-major: codegen.jl:116-118)
    want = {"BlackScholes": (["r", "σ"], 1, 1), "Heston": (["r", "κ", "θ", "σ", "ρ"], 2, 2),
            "TestModel": (["a", "c", "e", "b", "d", "f"], 3, 4), "Wiener": ([], 1, 1), "Deterministic": ([], 1, 0),
            "TimeDependent": (["a", "b"], 1, 1), "CoxIngersollRoss": (["κ", "θ", "σ"], 1, 1),
            "OrnsteinUhlenbeck": (["θ", "μ", "σ"], 1, 1), "FitzHughNagumo": (["ɛ", "s", "γ", "β", "σ1", "σ2"], 2, 2),
            "Lorenz": (["σ", "ρ", "β", "σ1", "σ2", "σ3"], 3, 3), "SpecialCaseModel1": ([], 1, 2),
            "SpecialCaseModel2": ([], 2, 1), "SpecialCaseModel3": ([], 3, 2)}
    for name, (params, D, M) in want.items():
        m = twin_model(name)
        assert (m.params, m.D, m.M) == (params, D, M), name


def test_twin_parameter_order_equals_the_front_end(S):
    for name, *_ in CASES:
        cls = S.sde_model(name + "_tw", MODELS.get(name) or REGISTRY[name])
        assert list(S.fieldnames(cls)) == twin_model(name).params, name


def test_reference_known_answers_through_the_twin():
    # test/runtests.jl:65-71 (drift), :77-81 (diffusion incl. the Heston Cholesky row), :89, :91 (noise-free trajectories)
    bs, he, tm = twin_model("BlackScholes"), twin_model("Heston"), twin_model("TestModel")
    assert bs.drift((0.01, 0.3), 0.0, [100.0]) == [100.0 * 0.01]
    assert he.drift((0.01, 10.0, 0.3, 0.1, 0.4), 0.0, [100.0, 0.4]) == [0.01 * 100.0, 10.0 * (0.3 - 0.4)]
    assert tm.drift((1, 2, 3, 4, 5, 6), 0.0, [0.1, 0.2, 0.3]) == [0.2 * 1, 0.3 * 2, 0.1 * 3]
    assert np.isclose(twin_model("TimeDependent").drift((1, 2), 4.0, [3.0])[0], 12)
    s = np.sqrt(0.4)
    assert np.allclose(he.diffusion_values((0.01, 10.0, 0.3, 0.1, 0.4), 0.0, [100.0, 0.4]),
                       [[100 * s, 0], [0.4 * 0.1 * s, np.sqrt(1 - 0.4 ** 2) * 0.1 * s]])
    assert tm.diffusion_values((1, 2, 3, 4, 5, 6), 0.0, [0.1, 0.2, 0.3]) == [[4, 5, 0, 0], [0, 1, 1, 0], [0, 0, 6, -1]]
    det = twin_model("Deterministic")
    assert T.sample_z(det, T.EM, (), 0.0, 1.0, [1.0], [[[0.0]]])[0][0] == 2.0
    x = T.simulate_wiener(det, T.EM, (), 0.0, 1.0, [1.0], [[[0.0]] * 6])[0]
    assert [r[0] for r in x] == [1, 2, 4, 8, 16, 32]


@pytest.mark.parametrize("name,params,x,t", CASES)
def test_functors_and_steps_bitwise(O, name, params, x, t):
    m = twin_model(name)
    D, M, _ = O.dims(name)
    rng = np.random.default_rng(7)
    for trial in range(8):
        xv = list(np.asarray(x, dtype=np.float64) * (1 + 0.2 * rng.random(D)))
        tt = t + 0.37 * trial
        assert m.drift(params, tt, xv) == list(O.drift(name, params, tt, xv)), name
        sg = O.diffusion(name, params, tt, xv)
        assert m.diffusion_values(params, tt, xv) == [list(r) for r in sg], name
        dt = 2.0 ** -rng.integers(3, 11)
        dw = list(np.sqrt(dt) * rng.standard_normal(max(M, 1)))
        for sid in [T.EM] + ([T.MILSTEIN, T.RUNGE_KUTTA] if M == 1 else []):
            assert T.step(m, sid, params, dt, tt, xv, dw) == list(O.step(name, sid, params, dt, tt, xv, dw)), (name, sid)


def test_dual_number_derivatives_equal_the_hand_coded_ones(O):
    """the C oracle hand-codes ∂σ (sde_oracle.c: bs_ddiff, cir_ddiff, td_ddiff, ...); the twin differentiates the model text"""
    assert twin_model("BlackScholes").diffusion_jacobian((0.01, 0.3), 0.0, [123.0]) == [[0.3]]
    cir = twin_model("CoxIngersollRoss").diffusion_jacobian((1.3, 0.04, 0.2), 0.0, [0.0625])
    assert cir == [[0.2 * (0.5 / np.sqrt(0.0625))]]
    assert twin_model("TimeDependent").diffusion_jacobian((1.0, 2.0), 3.0, [5.0]) == [[2.0]]
    assert twin_model("OrnsteinUhlenbeck").diffusion_jacobian((2.0, 0.7, 0.3), 0.0, [1.0]) == [[0.0]]
    assert twin_model("SpecialCaseModel2").diffusion_jacobian((), 0.0, [1.0, 2.0]) == [[0.0, 0.0], [0.0, 0.0]]


@pytest.mark.parametrize("name,params,x,t", CASES)
def test_drivers_bitwise(O, name, params, x, t):
    """sample, Wiener-driven simulate! and the coupled level sampler: twin == C oracle, bit for bit"""
    m = twin_model(name)
    D, M, _ = O.dims(name)
    MW = max(M, 1)
    npaths, nsteps, dt = 5, 24, 1.0 / 64
    z = np.random.default_rng(11).standard_normal((npaths, nsteps, MW))
    if name == "CoxIngersollRoss":
        params, x = (1.3, 0.09, 0.1), [0.09]      # away from the r = 0 boundary
    for sid in [T.EM] + ([T.MILSTEIN, T.RUNGE_KUTTA] if M == 1 else []):
        ref = O.sample_z(name, sid, params, t, dt, x, z)
        assert T.sample_z(m, sid, params, t, dt, list(map(float, x)), z.tolist()) == ref.tolist(), (name, sid)
        w = O.wiener_z(dt, z)
        ref = O.simulate_wiener(name, sid, params, t, dt, x, w)
        assert T.simulate_wiener(m, sid, params, t, dt, list(map(float, x)), w.tolist()) == ref.tolist(), (name, sid)
        if D == MW and M > 0:
            for sub in (2, 3, 4):
                zz = z[:, : (nsteps // sub) * sub]
                rc, rf = O.multilevel_sample_z(name, sid, params, t, dt * sub, dt, sub, x, zz)
                tc, tf = T.multilevel_sample_z(m, sid, params, t, dt * sub, dt, sub, list(map(float, x)), zz.tolist())
                assert tc == rc.tolist() and tf == rf.tolist(), (name, sid, sub)
