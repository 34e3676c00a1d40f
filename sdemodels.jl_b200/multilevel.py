"""Multilevel Monte Carlo: MultilevelScheme, level schemes, the budget allocator, and the coupled
fine/coarse level drivers (src/multilevel.jl).  The level loop lives here (host logic, as in the reference);
each level pair is one fused kernel behind `sdeb200_mlmc_*`."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .schemes import subdivide

_PD = C.POINTER(C.c_double)


@dataclass(frozen=True)
class MultilevelScheme:
    """struct MultilevelScheme{T}; scheme::T; substeps::Int64; levels::UnitRange{Int64}   (multilevel.jl:1-5)"""
    scheme: object
    substeps: int
    levels: range

    def __post_init__(self):
        lv = self.levels
        if not isinstance(lv, range):
            lv = range(int(lv[0]), int(lv[-1]) + 1)
            object.__setattr__(self, "levels", lv)
        if lv.step != 1 or len(lv) < 1:
            raise ValueError("levels must be a unit range, e.g. range(0, 5) for Julia's 0:4")

    @property
    def dt(self):
        return self.scheme.dt


def level_stream(stream, i):
    """SDEB200_LEVEL_STREAM (include/sdeb200.h): Philox stream word of level pair i of a multilevel run on `stream` — the
    pair index sits in the top byte, so a plain run on stream 1 does not share its noise with level 1 (pair 0 IS the plain
    `sample` on the caller's stream, multilevel.jl:56)."""
    if not 0 <= int(stream) < (1 << 24):
        raise ValueError("multilevel runs need 0 <= stream < 2^24")
    return int(stream) + (int(i) << 24)


def schemes(s):
    """[subdivide(s.scheme, s.substeps^l) for l in s.levels]   (multilevel.jl:7-10)"""
    return [subdivide(s.scheme, s.substeps ** l) for l in s.levels]


def npaths(s, max_budget):
    """npaths(s::MultilevelScheme, max_budget)   (multilevel.jl:12-24): equal-cost allocation, integer logic
    reproduced operation for operation (floor of a Float64 quotient, integer div, dot)."""
    budget = max_budget
    nlevels = len(s.levels)
    steps = [s.substeps ** l for l in s.levels]
    n_total = [0] * nlevels
    for i in range(nlevels, 0, -1):
        level_budget = min(int(np.floor(budget / i / steps[j])) * steps[j] for j in range(i))
        n_partial = [level_budget // steps[j] for j in range(i)]
        for j in range(i):
            n_total[j] += n_partial[j]
        budget -= sum(n_partial[j] * steps[j] for j in range(i))
    return np.array(n_total, dtype=np.int64)


def cost(s, n):
    """cost(s, npaths) = substeps.^levels .* npaths   (multilevel.jl:26-29)"""
    return np.array([s.substeps ** l for l in s.levels], dtype=np.int64) * np.asarray(n, dtype=np.int64)


def _level_counts(s, n):
    if np.isscalar(n) or n is None:
        n = 1 if n is None else n
        return [int(n)] * len(s.levels)  # fill(npaths, length(levels))   (multilevel.jl:31-35)
    n = [int(v) for v in n]
    if len(n) != len(s.levels):
        raise ValueError("npaths must have one entry per level")
    return n


def _ml_sample(model, s, t0, x0, nsteps, n, *, seed, precision, math, stream, path_offset, device, rng=None):
    """sample(model, ::MultilevelScheme, ...) -> [x_L0, xc_1, xf_1, ...]   (multilevel.jl:52-63, 104-108)"""
    from . import simulation as sim
    L = _lib.lib()
    counts = _level_counts(s, n)
    ml_schemes = schemes(s)
    ml_steps = [s.substeps ** l for l in s.levels]
    v0, scalar = sim._x0(model, x0)
    p = sim._params(model)
    out = [sim.sample(model, ml_schemes[0], t0, x0, nsteps * ml_steps[0], counts[0], seed=seed, precision=precision,
                      math=math, stream=stream, path_offset=path_offset, rng=rng, device=device)]
    for i in range(1, len(s.levels)):
        coarse, fine = ml_schemes[i - 1], ml_schemes[i]
        shape = (counts[i],) if scalar else (counts[i], model._D)
        xc, xf = sim._Buf(shape, precision, None, device), sim._Buf(shape, precision, None, device)
        o = sim._opts(coarse, t0, seed, precision, math, level_stream(stream, i), path_offset, rng=rng)
        _lib.check(L.sdeb200_mlmc_sample_level(model._handle, C.byref(o), p.ctypes.data_as(_PD),
                                               v0.ctypes.data_as(_PD), fine.dt, s.substeps,
                                               nsteps * ml_steps[i - 1], counts[i], xc.ptr, xf.ptr))
        out += [xc.arr, xf.arr]
    return out


def multilevel_sample_level(model, coarse, fine, substeps, t0, x0, nsteps, n, *, seed=0, precision="f64",
                            math="fast", stream=0, path_offset=0, rng=None, device=False):
    """_multilevel_sample(model, coarse_scheme, fine_scheme, substeps, t0, x0, nsteps, npaths) -> (xc, xf)
    (multilevel.jl:65-89); nsteps counts COARSE steps."""
    from . import simulation as sim
    L = _lib.lib()
    v0, scalar = sim._x0(model, x0)
    p = sim._params(model)
    shape = (n,) if scalar else (n, model._D)
    xc, xf = sim._Buf(shape, precision, None, device), sim._Buf(shape, precision, None, device)
    o = sim._opts(coarse, t0, seed, precision, math, stream, path_offset, rng=rng)
    _lib.check(L.sdeb200_mlmc_sample_level(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                           fine.dt, substeps, nsteps, n, xc.ptr, xf.ptr))
    return xc.arr, xf.arr


def multilevel_simulate_level(model, coarse, fine, substeps, t0, x0, nsteps, n, *, seed=0, precision="f64",
                              math="fast", stream=0, path_offset=0, rng=None, device=False):
    """_multilevel_simulate(...) -> (xc, xf, tc, tf)   (multilevel.jl:91-100)"""
    from . import simulation as sim
    L = _lib.lib()
    v0, scalar = sim._x0(model, x0)
    p = sim._params(model)
    nf, nc = nsteps * substeps + 1, nsteps + 1
    fs = (n, nf) if scalar else (n, nf, model._D)
    cs = (n, nc) if scalar else (n, nc, model._D)
    xc, xf = sim._Buf(cs, precision, None, device), sim._Buf(fs, precision, None, device)
    o = sim._opts(coarse, t0, seed, precision, math, stream, path_offset, rng=rng)
    _lib.check(L.sdeb200_mlmc_simulate_level(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                             fine.dt, substeps, nsteps, n, xc.ptr, xf.ptr))
    tf = t0 + fine.dt * np.arange(nsteps * substeps + 1)   # multilevel.jl:92
    tc = tf[::substeps]                                      # multilevel.jl:93
    return xc.arr, xf.arr, tc, tf


def _ml_simulate(model, s, t0, x0, nsteps, n, *, seed, precision, math, stream, path_offset, device, rng=None):
    """simulate(model, ::MultilevelScheme, ...) -> (x, t), 2L-1 entries each   (multilevel.jl:37-50, 110-115)"""
    from . import simulation as sim
    counts = _level_counts(s, n)
    ml_schemes = schemes(s)
    ml_steps = [s.substeps ** l for l in s.levels]
    x1, t1 = sim.simulate(model, ml_schemes[0], t0, x0, nsteps * ml_steps[0], counts[0], seed=seed,
                          precision=precision, math=math, stream=stream, path_offset=path_offset, rng=rng, device=device)
    xs, ts = [x1], [t1]
    for i in range(1, len(s.levels)):
        xc, xf, tc, tf = multilevel_simulate_level(model, ml_schemes[i - 1], ml_schemes[i], s.substeps, t0, x0,
                                                   nsteps * ml_steps[i - 1], counts[i], seed=seed,
                                                   precision=precision, math=math, stream=level_stream(stream, i),
                                                   path_offset=path_offset, rng=rng, device=device)
        xs += [xc, xf]
        ts += [tc, tf]
    return xs, ts


@dataclass
class MultilevelEstimate:
    mean: float             # Σ_l mean(Y_l)
    stderr: float           # sqrt(Σ_l var(Y_l)/n_l)
    level_mean: np.ndarray  # mean of Y_l = P(xf_l) - P(xc_l)   (level 0: P(x_0))
    level_var: np.ndarray
    fine_mean: np.ndarray   # mean of P(xf_l)
    n: np.ndarray
    nonfinite: np.ndarray
    sums: np.ndarray        # [L, 6] exactly accumulated ΣY, ΣY², ΣPf, ΣPf², ΣPc, ΣPc²


def ml_estimate(model, s, payoff, t0, x0, nsteps, n, *, seed=0, precision="f64", math="fast", stream=0,
                path_offset=0, rng=None, level_offsets=None):
    """Multilevel estimator of E[payoff(X_T)]: all levels run concurrently on separate streams, each level ends
    in its own all-reduce.  `n`: this rank's paths per level."""
    from . import simulation as sim
    L = _lib.lib()
    counts = np.array(_level_counts(s, n), dtype=np.int64)
    v0, _ = sim._x0(model, x0)
    p = sim._params(model)
    nl = len(s.levels)
    res = np.zeros((nl, 8))
    o = sim._opts(s.scheme, t0, seed, precision, math, stream, path_offset, rng=rng)
    pay = sim._payoff(payoff)
    # level_offsets: this rank's first path of every level (multi-process runs: every level sharded over all ranks)
    offs = None if level_offsets is None else np.ascontiguousarray(level_offsets, dtype=np.int64)
    _lib.check(L.sdeb200_mlmc_estimate_sharded(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                               s.substeps, s.levels[0], nl, int(nsteps),
                                               counts.ctypes.data_as(C.POINTER(C.c_int64)),
                                               None if offs is None else offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                               C.byref(pay), res.ctypes.data_as(_PD)))
    ntot, bad = res[:, 6], res[:, 7]
    good = np.maximum(ntot - bad, 1)
    m = res[:, 0] / good
    var = np.maximum(res[:, 1] / good - m * m, 0.0) * good / np.maximum(good - 1, 1)
    return MultilevelEstimate(float(m.sum()), float(np.sqrt((var / good).sum())), m, var, res[:, 2] / good,
                              ntot.astype(np.int64), bad.astype(np.int64), res[:, :6].copy())


@dataclass
class AdaptiveMLMC:
    mean: float
    stderr: float            # sqrt(sum_l V_l / N_l)
    bias_estimate: float     # |E[Y_L]| / (substeps^alpha - 1), alpha = 1 (weak order of the schemes on the path)
    levels: range
    n: np.ndarray            # paths used per level
    level_mean: np.ndarray
    level_var: np.ndarray
    cost: np.ndarray         # fine + coarse steps per path, per level
    iterations: int


def mlmc_adaptive(model, scheme, substeps, payoff, t0, x0, nsteps, eps, *, lmin=2, lmax=10, n0=1 << 14, seed=0,
                  precision="f64", math="fast", stream=0, alpha=1.0, rng=None):
    """Giles' adaptive multilevel estimator on top of the per-level exact sums (the reference
    stops at the raw level samples, multilevel.jl:52-63).  Levels 0..L with base `scheme`; N_l is raised to
    ceil(2 eps^-2 sqrt(V_l / C_l) sum_k sqrt(V_k C_k)); a level is added while the extrapolated bias
    |E[Y_L]| / (substeps^alpha - 1) exceeds eps / sqrt(2).  Additional paths of a level CONTINUE its path
    numbering (path_offset = paths already done), so the final sums equal those of one run with the final N_l."""
    from . import simulation as sim
    L_ = _lib.lib()
    v0, _ = sim._x0(model, x0)
    p = sim._params(model)
    pay = sim._payoff(payoff)
    sums = np.zeros((lmax + 1, 2))
    done = np.zeros(lmax + 1, dtype=np.int64)

    def run(level, n_add):
        if n_add <= 0:
            return
        if level == 0:
            e = sim.estimate(model, scheme, payoff, t0, x0, nsteps, int(n_add), seed=seed, precision=precision,
                             math=math, stream=stream, path_offset=int(done[0]), rng=rng)
            sums[0] += e.sums[0]
        else:
            coarse = subdivide(scheme, substeps ** (level - 1))
            fine = subdivide(scheme, substeps ** level)
            o = sim._opts(coarse, t0, seed, precision, math, level_stream(stream, level), int(done[level]), rng=rng)
            out = np.zeros(8)
            _lib.check(L_.sdeb200_mlmc_estimate_level(model._handle, C.byref(o), p.ctypes.data_as(_PD),
                                                      v0.ctypes.data_as(_PD), fine.dt, substeps,
                                                      nsteps * substeps ** (level - 1), int(n_add), C.byref(pay),
                                                      out.ctypes.data_as(_PD)))
            sums[level] += out[:2]
        done[level] += n_add

    L = lmin
    for l in range(L + 1):
        run(l, n0)
    it = 0
    while True:
        it += 1
        n = done[:L + 1].astype(np.float64)
        m = sums[:L + 1, 0] / n
        var = np.maximum(sums[:L + 1, 1] / n - m * m, 1e-300)
        cost = np.array([nsteps * substeps ** l * (1.0 + (1.0 / substeps if l else 0.0)) for l in range(L + 1)])
        want = np.ceil(2.0 / eps ** 2 * np.sqrt(var / cost) * np.sum(np.sqrt(var * cost))).astype(np.int64)
        extra = np.maximum(want - done[:L + 1], 0)
        if np.any(extra > 0.01 * done[:L + 1]):
            for l in range(L + 1):
                run(l, int(extra[l]))
            continue
        bias = abs(m[L]) / (substeps ** alpha - 1.0)
        if bias > eps / np.sqrt(2.0) and L < lmax:
            L += 1
            run(L, n0)
            continue
        break
    n = done[:L + 1].astype(np.float64)
    m = sums[:L + 1, 0] / n
    var = np.maximum(sums[:L + 1, 1] / n - m * m, 0.0)
    return AdaptiveMLMC(float(m.sum()), float(np.sqrt((var / n).sum())), float(abs(m[L]) / (substeps ** alpha - 1.0)),
                        range(0, L + 1), done[:L + 1].copy(), m, var, cost, it)
