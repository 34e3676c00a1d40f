"""The @sde_model front end (csrc/sde_dsl.cpp) against the reference's codegen / model tests
(test/runtests.jl:7-82) and against the hand-written oracle functors.  Runs without a GPU: expressions are
checked with the front end's host evaluator (test infrastructure) and the generated CUDA functors are compiled
for sm_100a with NVRTC (compilation needs no device)."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import BS_P, HESTON_P

# test/runtests.jl:17-41 + registry models (src/models/models.jl, black_scholes.jl, cox_ingersoll_ross.jl, ornstein_uhlenbeck.jl)
MODELS = {
    "TestModel": "dX = Y*a*dt + b*dW1 + d*dW2\ndY = Z*c*dt +   dW2 +   dW3\ndZ = X*e*dt + f*dW3 -   dW4",
    "Wiener": "dX = dW",
    "Deterministic": "dX = X*dt",
    "TimeDependent": "dX = t*a*X*dt + b*X*dW",
    "SpecialCaseModel1": "dx = dw1 + dw2",
    "SpecialCaseModel2": "dx = dw1\ndy = dw1",
    "SpecialCaseModel3": "dx = dw1\ndy = dw1 + dw2\ndz = dw2",
}


def build(S, name):
    eqs = MODELS.get(name) or S.REGISTRY[name]
    return S.sde_model(name + "_t", eqs)


def exprs(S, cls):
    L = S._lib.lib()
    D, M = cls._D, max(cls._M, 1)
    dr = [L.sdeb200_model_drift_expr(cls._handle, i).decode() for i in range(D)]
    df = [[L.sdeb200_model_diffusion_expr(cls._handle, i, j).decode() for j in range(M)] for i in range(D)]
    return dr, df


def heval(S, expr, names, values):
    L = S._lib.lib()
    fn = L.sdeb200_debug_eval
    fn.restype = C.c_int
    out = C.c_double()
    arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
    vals = (C.c_double * max(len(values), 1))(*values)
    rc = fn(expr.encode(), arr, vals, len(names), C.byref(out))
    assert rc == 0, f"cannot evaluate {expr}"
    return out.value


def test_model_creation(S):
    # runtests.jl:50-59
    assert S.fieldnames(build(S, "BlackScholes")) == ("r", "σ")
    assert S.fieldnames(build(S, "Heston")) == ("r", "κ", "θ", "σ", "ρ")
    assert S.fieldnames(build(S, "TestModel")) == ("a", "c", "e", "b", "d", "f")
    assert S.dim(build(S, "BlackScholes")(0.01, 0.3)) == (1, 1)
    assert S.dim(build(S, "Heston")(*HESTON_P)) == (2, 2)
    assert S.dim(build(S, "TestModel")(1, 2, 3, 4, 5, 6)) == (3, 4)
    assert S.dim(build(S, "Wiener")()) == (1, 1)
    assert S.dim(build(S, "Deterministic")()) == (1, 0)
    assert S.variables(build(S, "Heston")(*HESTON_P)) == ("S", "V")
    # registry models: parameters in order of appearance
    assert S.fieldnames(build(S, "CoxIngersollRoss")) == ("κ", "θ", "σ")
    assert S.fieldnames(build(S, "OrnsteinUhlenbeck")) == ("θ", "μ", "σ")
    assert S.fieldnames(build(S, "FitzHughNagumo")) == ("ɛ", "s", "γ", "β", "σ1", "σ2")
    assert S.fieldnames(build(S, "Lorenz")) == ("σ", "ρ", "β", "σ1", "σ2", "σ3")
    assert S.dim(build(S, "SpecialCaseModel1")()) == (1, 2)
    assert S.dim(build(S, "SpecialCaseModel2")()) == (2, 1)
    assert S.dim(build(S, "SpecialCaseModel3")()) == (3, 2)
    # built-in handles agree with the generated ones
    assert S.fieldnames(S.BlackScholes) == ("r", "σ") and S.fieldnames(S.Heston) == ("r", "κ", "θ", "σ", "ρ")


def test_model_kind(S):
    # codegen.jl:120-126
    assert build(S, "Heston")._kind == 0            # AbstractSDE
    assert build(S, "Wiener")._kind == 1            # StateIndependentDiffusion
    assert build(S, "OrnsteinUhlenbeck")._kind == 1
    assert build(S, "TestModel")._kind == 1
    with pytest.raises(S.SDEB200Error) as ei:       # JumpSDE: outside the accelerated path
        S.sde_model("Bates", "dS = α*S*dt + sqrt(V)*S*dW1 + S*(ξ-1.0)*dN\ndV = κ*(θ-V)*dt + σ*sqrt(V)*dW2\n"
                             "dW1*dW2 = ρ*dt\nξ ~ LogNormal(μ, δ)", "α", "κ", "θ", "σ", "ρ", "λ", "μ", "δ")
    assert ei.value.code == S._lib.E_UNSUPPORTED


def test_heston_generated_form(S):
    """The generated expressions are the ones the macro produces (README.md:42-59 / SURVEY §3.5): Cholesky of
    the correlation matrix folded into the diffusion, structural zero kept."""
    dr, df = exprs(S, build(S, "Heston"))
    assert dr == ["r * S", "κ * (θ - V)"]
    assert df == [["sqrt(V) * S", "0"], ["σ * sqrt(V) * ρ", "σ * sqrt(V) * sqrt(1 - ρ ^ 2)"]]
    dr, df = exprs(S, build(S, "TestModel"))
    assert dr == ["Y * a", "Z * c", "X * e"]
    assert df == [["b", "d", "0", "0"], ["0", "1", "1", "0"], ["0", "0", "f", "-1"]]
    dr, df = exprs(S, build(S, "Deterministic"))
    assert dr == ["X"] and df == [["0"]]


CASES = [
    ("BlackScholes", BS_P, [100.0], 0.0),
    ("Heston", HESTON_P, [100.0, 0.4], 0.0),
    ("TestModel", (1, 2, 3, 4, 5, 6), [0.1, 0.2, 0.3], 0.0),
    ("Wiener", (), [100.0], 0.0),
    ("Deterministic", (), [100.0], 0.0),
    ("TimeDependent", (1, 2), [3.0], 4.0),
    ("SpecialCaseModel1", (), [100.0], 0.0),
    ("SpecialCaseModel2", (), [100.0, 100.0], 0.0),
    ("SpecialCaseModel3", (), [100.0, 100.0, 100.0], 0.0),
    ("CoxIngersollRoss", (1.3, 0.04, 0.2), [0.05], 0.0),
    ("OrnsteinUhlenbeck", (2.0, 0.7, 0.3), [1.1], 0.0),
    ("FitzHughNagumo", (10.0, 0.1, 1.5, 0.8, 0.3, 0.4), [0.3, -0.2], 0.0),
    ("Lorenz", (10.0, 28.0, 8 / 3, 0.1, 0.2, 0.3), [1.0, 2.0, 3.0], 0.0),
]


@pytest.mark.parametrize("name,params,x,t", CASES)
def test_drift_and_diffusion_equal_the_oracle_functors(S, O, name, params, x, t):
    """The reference's drift / diffusion known answers (runtests.jl:61-82) hold for the oracle
    (tests/test_oracle_kat.py); here the front end's expressions must evaluate to the oracle's values with the
    same rounding sequence (bit-identical)."""
    cls = build(S, name)
    dr, df = exprs(S, cls)
    names = list(cls._fields) + list(cls._vars) + ["t"]
    rng = np.random.default_rng(1)
    for trial in range(3):
        xv = list(np.asarray(x) * (1 + 0.1 * trial * rng.random(len(x))))
        vals = list(map(float, params)) + xv + [t + trial]
        mu = O.drift(name, params, t + trial, xv)
        sg = O.diffusion(name, params, t + trial, xv)
        for i in range(cls._D):
            assert heval(S, dr[i], names, vals) == mu[i]
            for j in range(max(cls._M, 1)):
                assert heval(S, df[i][j], names, vals) == sg[i, j]


def test_reference_drift_diffusion_known_answers(S):
    # runtests.jl:65-71,77-81 directly on the front end's expressions
    H = build(S, "Heston")
    dr, df = exprs(S, H)
    names, vals = list(H._fields) + ["S", "V", "t"], list(HESTON_P) + [100.0, 0.4, 0.0]
    assert np.isclose(heval(S, dr[0], names, vals), 0.01 * 100.0)
    assert np.isclose(heval(S, dr[1], names, vals), 10.0 * (0.3 - 0.4))
    want = [[100.0 * math.sqrt(0.4), 0.0], [0.4 * 0.1 * math.sqrt(0.4), math.sqrt(1 - 0.4 ** 2) * 0.1 * math.sqrt(0.4)]]
    for i in range(2):
        for j in range(2):
            assert np.isclose(heval(S, df[i][j], names, vals), want[i][j])
    T = build(S, "TimeDependent")
    dr, _ = exprs(S, T)
    assert np.isclose(heval(S, dr[0], ["a", "b", "X", "t"], [1, 2, 3.0, 4.0]), 12)


def test_simplifier_rules(S):
    """Calculus.simplify's rule set as the macro relies on it (codegen.jl:178-183,213-237)."""
    L = S._lib.lib()

    def simp(e):
        buf = C.create_string_buffer(512)
        assert L.sdeb200_debug_simplify(e.encode(), buf, 512) == 0
        return buf.value.decode()
    assert simp("Y*a*1 + b*0 + d*0") == "Y * a"
    assert simp("Z*c*0 + 1 + 0") == "1"
    assert simp("(X*e*0 + f*0) - 1") == "-1"
    assert simp("ρ - (0 + 1*ρ)") == "0"
    assert simp("(ρ - (0 + 1*ρ)) / sqrt(1 - ρ^2)") == "0"
    assert simp("1 - (0 + ρ^2)") == "1 - ρ ^ 2"
    assert simp("x^1 + y^0") == "1 + x"          # numeric arguments are summed and moved to the front
    assert simp("2*x*3") == "6 * x"
    assert simp("x/1") == "x" and simp("0/x") == "0" and simp("x - 0") == "x" and simp("0 - x") == "-x"
    assert simp("2θ*x") == "2 * θ * x"           # numeric-literal juxtaposition


def test_correlated_three_factor_cholesky(S):
    """A 3x3 correlation matrix: σ·chol(R) reproduces R as the instantaneous covariance (codegen.jl:221-237)."""
    eqs = ("dx = a*dw1\ndy = b*dw2\ndz = c*dw3\ndw1*dw2 = r12*dt\ndw1*dw3 = r13*dt\ndw2*dw3 = r23*dt")
    cls = S.sde_model("Corr3", eqs)
    assert cls._fields == ("a", "b", "r12", "c", "r13", "r23") or set(cls._fields) == {"a", "b", "c", "r12", "r13", "r23"}
    _, df = exprs(S, cls)
    p = dict(a=1.5, b=0.7, c=2.0, r12=0.3, r13=-0.2, r23=0.5)
    names = list(cls._fields) + ["x", "y", "z", "t"]
    vals = [p[k] for k in cls._fields] + [0.0, 0.0, 0.0, 0.0]
    Sg = np.array([[heval(S, df[i][j], names, vals) for j in range(3)] for i in range(3)])
    R = np.array([[1, 0.3, -0.2], [0.3, 1, 0.5], [-0.2, 0.5, 1]])
    sd = np.diag([1.5, 0.7, 2.0])
    assert np.allclose(Sg @ Sg.T, sd @ R @ sd, atol=1e-14)
    assert np.allclose(np.triu(Sg, 1), 0)


def test_explicit_parameter_order_and_errors(S):
    cls = S.sde_model("BSx", "dS = r*S*dt + σ*S*dW", "σ", "r")
    assert cls._fields == ("σ", "r")
    with pytest.raises(S.SDEB200Error) as ei:
        S.sde_model("Bad", "dS = r*S*dt + q*S*dW", "r")  # q unbound
    assert ei.value.code == S._lib.E_PARSE
    with pytest.raises(S.SDEB200Error) as ei:
        S.sde_model("Bad2", "this is not an equation")
    assert ei.value.code == S._lib.E_PARSE
    with pytest.raises(S.SDEB200Error):
        S.sde_model("Bad3", "dS = r*S*dt + (σ*S*dW")


@pytest.mark.parametrize("name", ["Heston", "BlackScholes", "TestModel", "Deterministic", "TimeDependent",
                                  "SpecialCaseModel2", "SpecialCaseModel3", "CoxIngersollRoss", "FitzHughNagumo", "Lorenz"])
def test_generated_functor_compiles_for_sm100a(S, name):
    """NVRTC builds every path kernel around the generated functor (no device needed to compile)."""
    L = S._lib.lib()
    cls = build(S, name)
    src = L.sdeb200_model_cuda_source(cls._handle)
    fn = L.sdeb200_debug_nvrtc_check
    fn.restype = C.c_int
    log = C.create_string_buffer(1 << 16)
    variants = [(0, 0)] if name not in ("Heston", "BlackScholes") else [(0, 0), (0, 1), (1, 0)]
    if name in ("Lorenz", "TestModel"):
        variants = [(0, 0), (1, 0)]   # D = 3: 24- and 12-byte rows, which the TMA kernel forms must compile around (as no-ops)
    for prec, math_ in variants:
        rc = fn(src, cls._M, prec, math_, log, len(log))
        assert rc > 1000, log.value.decode()


@pytest.mark.parametrize("name", ["BlackScholes", "Heston"])
def test_every_kernel_family_builds_under_nvrtc(S, name):
    """NVRTC (JIT mode) is stricter than nvcc: every kernel family — terminal, RNG-driven trajectories (incl. the TMA
    store kernel), Wiener-driven, coefficient / logtransition — of every slot (f64 stream v1, f32, f64 stream v2) must
    compile from the generated functor.  M = 1 covers EM / Milstein / RungeKutta, M = 2 the two-noise batches."""
    L = S._lib.lib()
    cls = build(S, name)
    src = L.sdeb200_model_cuda_source(cls._handle)
    fn = L.sdeb200_debug_nvrtc_check_family
    fn.restype = C.c_int
    log = C.create_string_buffer(1 << 16)
    for slot in (0, 1, 2):
        for family in (0, 1, 2, 3):
            if slot == 2 and family in (2, 3) and name == "Heston":
                continue     # no noise in these families: stream v2 compiles the same code as slot 0
            rc = fn(src, cls._M, slot, 0, family, log, len(log))
            assert rc > 1000, (slot, family, log.value.decode()[:2000])


def test_cubin_cache_round_trip(S, tmp_path, monkeypatch):
    """A generated model's cubin is written to $SDEB200_CACHE_DIR and loaded back by the next compile of the same
    source / options; an empty variable turns the cache off; a corrupt entry is ignored."""
    import time
    L = S._lib.lib()
    cls = S.sde_model("CacheProbe", "dX = a*(b - X)*dt + c*X*dW1; dY = e*Y*dt + f*dW2")   # M = 2: one scheme to compile
    L.sdeb200_model_cuda_source.restype = C.c_char_p
    src = L.sdeb200_model_cuda_source(cls._handle)
    fn = L.sdeb200_debug_nvrtc_check
    fn.restype = C.c_int
    log = C.create_string_buffer(1 << 16)
    monkeypatch.setenv("SDEB200_CACHE_DIR", str(tmp_path / "cubins"))
    t = time.perf_counter()
    n1 = fn(src, 2, 1, 0, log, len(log))
    cold = time.perf_counter() - t
    files = list((tmp_path / "cubins").glob("*.cubin"))   # one cubin per kernel family compiled (terminal + coefficient/logtransition)
    assert n1 > 1000 and len(files) == 2 and sum(f.stat().st_size for f in files) == n1
    t = time.perf_counter()
    n2 = fn(src, 2, 1, 0, log, len(log))
    warm = time.perf_counter() - t
    assert n2 == n1 and warm < 0.5 * cold
    files[0].write_bytes(b"not a cubin" * 10)                 # corrupt entry: recompiled and replaced
    assert fn(src, 2, 1, 0, log, len(log)) == n1 and sum(f.stat().st_size for f in files) == n1
    monkeypatch.setenv("SDEB200_CACHE_DIR", "")               # off
    assert fn(src, 2, 1, 1, log, len(log)) > 1000 and len(list((tmp_path / "cubins").glob("*.cubin"))) == 2


def test_polar_step_is_emitted_for_a_common_sqrt_factor(S):
    """M == 2 models whose non-zero diffusion entries all carry the same state-dependent factor sqrt(A) get
    GenModel::step_polar (one square root of A*w per step instead of the radius' and the model's; the FP64 kernels take it
    in FAST math, sde_device.cuh Noise<>); every other model keeps the plain step only."""
    L = S._lib.lib()
    L.sdeb200_model_cuda_source.restype = C.c_char_p
    src = L.sdeb200_model_cuda_source(build(S, "Heston")._handle).decode()
    body = src[src.index("static constexpr int POLAR = 1;"):src.index("static void coefficients")]
    assert "const T G = SDE_SQRT(x1 * w);" in body and "(G * ca)" in body and "* ca) + " in body and "* sb)" in body
    assert "x[0] = ((x0 * (c[6] + (G * ca))) + x0);" in body    # geometric row: x0 factored out, r*dt folded into one constant
    assert "x[1] = ((x1 * c[9]) + (c[10] + (G * ((c[7] * ca) + (c[8] * sb)))));" in body   # mean reversion expanded: V*(1 - κΔt) + (κθΔt + noise)
    assert body.count("SDE_SQRT") == 1          # the only square root of the polar step
    three_halves = S.sde_model("ThreeHalves", "dS = r*S*dt + sqrt(V)*S*dW1; dV = κ*V*(θ-V)*dt + σ*V*sqrt(V)*dW2")
    src = L.sdeb200_model_cuda_source(three_halves._handle).decode()
    assert "POLAR" in src and "x[1] = ((x1 * ((c[5] * (c[2] - x1)) + (G * (c[3] * sb)))) + x1);" in src   # V factored out of κ*V*(θ-V)*dt and σ*V*sqrt(V)*dW2
    for name, text in (("MixedVol", "dS = r*S*dt + σ*S*dW1; dV = κ*(θ-V)*dt + ξ*sqrt(V)*dW2"),       # one entry without the factor
                       ("ParamRoot", "dS = r*S*dt + sqrt(a)*S*dW1; dV = κ*(θ-V)*dt + sqrt(a)*V*dW2"),  # parameter-only root: hoisted instead
                       ("TwoRoots", "dS = r*S*dt + sqrt(V)*S*dW1; dV = κ*(θ-V)*dt + sqrt(S)*dW2")):      # different arguments
        assert "POLAR" not in L.sdeb200_model_cuda_source(S.sde_model(name, text)._handle).decode(), name
    for name in ("BlackScholes", "CoxIngersollRoss", "Lorenz"):   # M == 1 / M == 3: never
        assert "POLAR" not in L.sdeb200_model_cuda_source(build(S, name)._handle).decode()


def test_fast_forms_of_the_generated_update(S):
    """FAST math (SDE_STRICT = 0) emits the Euler-Maruyama update with parameter-only factors and Δt folded into one
    constant per term and the state factored out of geometric rows; STRICT keeps the reference's literal order below it."""
    L = S._lib.lib()
    L.sdeb200_model_cuda_source.restype = C.c_char_p
    src = L.sdeb200_model_cuda_source(build(S, "BlackScholes")._handle).decode()
    fast = src[src.index("#if !SDE_STRICT"):src.index("#endif", src.index("#if !SDE_STRICT"))]
    assert "c[2] = (c[0] * dt);  // r * dt" in src and "x[0] = ((x0 * (c[2] + (c[1] * dw[0]))) + x0);" in fast
    assert "T y0 = (x0 + mu0 * dt) + (s0_0 * dw[0]);" in src          # the literal form: euler_maruyama.jl:4
    src = L.sdeb200_model_cuda_source(build(S, "CoxIngersollRoss")._handle).decode()
    assert "c[3] = ((T)1 + (((T)(-1) * c[0]) * dt));" in src and "c[4] = ((c[0] * dt) * c[1]);" in src
    assert "x[0] = ((x0 * c[3]) + (c[4] + ((c[2] * SDE_SQRT(x0)) * dw[0])));" in src   # mean reversion: x*(1 - κΔt) + (κθΔt + noise)
    # ... in Float64 only: FP32 keeps x + κΔt (θ - x), whose level θ does not move with the rounding of 1 - κΔt
    assert "} else {  // FP32: mean reversion stays x + c (a - x)" in src
    assert "x[0] = (x0 + ((c[5] * (c[1] - x0)) + ((c[2] * SDE_SQRT(x0)) * dw[0])));" in src
    src = L.sdeb200_model_cuda_source(build(S, "TimeDependent")._handle).decode()
    assert "(c[2] * t)" in src                                       # t stays a per-step factor
