"""Path drivers: sample / sample! / simulate / simulate! / wiener!  (src/simulation.jl:22-134).

Same call shapes as the reference; the loops, the noise and the `step` arithmetic run on the GPU behind the
C ABI.  Arrays come back in Julia's memory order, expressed in numpy's C order:

    Julia                                         numpy
    sample   -> Vector{Float64}(npaths)           (npaths,)                 scalar x0
             -> (D, npaths)                       (npaths, D)               vector x0
    simulate -> (nsteps+1, npaths)                (npaths, nsteps+1)        scalar x0
             -> (D, nsteps+1, npaths)             (npaths, nsteps+1, D)     vector x0

Backend keywords (no reference analogue): seed (Philox key, replaces Random.seed!), precision "f64"|"f32",
math "fast"|"strict", stream (independent sub-stream), path_offset (global index of the first path: shards of
one logical run), out= (numpy array or torch CUDA tensor to fill), device=True (return a torch CUDA tensor).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .multilevel import MultilevelScheme, _ml_sample, _ml_simulate

_PD = C.POINTER(C.c_double)
_PREC = {"f64": _lib.F64, "f32": _lib.F32, "float64": _lib.F64, "float32": _lib.F32}
_MATH = {"fast": _lib.MATH_FAST, "strict": _lib.MATH_STRICT}
DEFAULT_RNG = int(__import__("os").environ.get("SDEB200_RNG", "0"))   # sdeb200_opts.rng: 0/1 = normal stream v1 (the default), 2 = stream v2 (96 Philox bits per FP64 pair)


def _opts(scheme, t0, seed=0, precision="f64", math="fast", stream=0, path_offset=0, rng=None, dt=None):
    o = _lib.Opts()
    o.scheme = scheme._id
    o.precision = _PREC[precision]
    o.math = _MATH[math]
    o.t0 = float(t0)
    o.dt = float(scheme.dt if dt is None else dt)
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    o.stream = int(stream)
    o.path_offset = int(path_offset)
    o.rng = int(DEFAULT_RNG if rng is None else rng)
    return o


def _np_dtype(precision):
    return np.float64 if _PREC[precision] == _lib.F64 else np.float32


def _x0(model, x0):
    scalar = np.isscalar(x0)
    v = np.ascontiguousarray(np.atleast_1d(np.asarray(x0, dtype=np.float64)))
    if v.shape != (model._D,):
        raise ValueError(f"x0 must have {model._D} component(s) for {model._name}")
    return v, scalar


def _params(model):
    return model.params if len(model.params) else np.zeros(1)


class _Buf:
    """An output buffer: numpy (host) or torch CUDA tensor (device)."""

    def __init__(self, shape, precision, out=None, device=False):
        dt = _np_dtype(precision)
        self.torch = False
        if out is not None:
            if hasattr(out, "to_host"):  # DeviceBuffer (library-owned device memory)
                if out.dtype != dt or tuple(out.shape) != tuple(shape):
                    raise ValueError(f"out must be a {dt.__name__} DeviceBuffer of shape {tuple(shape)}")
                self.arr, self.ptr, self.torch = out, out.ptr, False
            elif hasattr(out, "data_ptr"):  # torch tensor
                import torch
                want = torch.float64 if dt == np.float64 else torch.float32
                if out.dtype != want or not out.is_contiguous() or tuple(out.shape) != tuple(shape):
                    raise ValueError(f"out must be a contiguous {want} tensor of shape {tuple(shape)}")
                self.arr, self.ptr, self.torch = out, out.data_ptr(), True
            else:
                if out.dtype != dt or not out.flags.c_contiguous or out.shape != tuple(shape):
                    raise ValueError(f"out must be a C-contiguous {dt.__name__} array of shape {tuple(shape)}")
                self.arr, self.ptr = out, out.ctypes.data
        elif device:
            import torch
            self.arr = torch.empty(tuple(shape), dtype=torch.float64 if dt == np.float64 else torch.float32, device="cuda")
            self.ptr, self.torch = self.arr.data_ptr(), True
        else:
            self.arr = np.empty(tuple(shape), dtype=dt)
            self.ptr = self.arr.ctypes.data


def sample(model, scheme, t0, x0, nsteps, npaths=None, *, seed=0, precision="f64", math="fast", stream=0,
           path_offset=0, rng=None, out=None, device=False):
    """sample(model, scheme, t0, x0, nsteps[, npaths])   (simulation.jl:39-48, 59-67; multilevel.jl:34-35,52-63)

    Terminal state(s) after `nsteps` steps of `scheme`.  Without `npaths`: one path (a float or a D-vector).
    With a MultilevelScheme: the list [x_L0, xc_1, xf_1, xc_2, xf_2, ...] of 2L-1 arrays."""
    if isinstance(scheme, MultilevelScheme):
        return _ml_sample(model, scheme, t0, x0, nsteps, npaths, seed=seed, precision=precision, math=math,
                          stream=stream, path_offset=path_offset, rng=rng, device=device)
    L = _lib.lib()
    v0, scalar = _x0(model, x0)
    single = npaths is None
    n = 1 if single else int(npaths)
    buf = _Buf((n,) if scalar else (n, model._D), precision, out, device)
    o = _opts(scheme, t0, seed, precision, math, stream, path_offset, rng=rng)
    p = _params(model)
    _lib.check(L.sdeb200_sample(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                int(nsteps), n, buf.ptr))
    if single:
        a = buf.arr.cpu().numpy() if buf.torch else buf.arr
        return float(a[0]) if scalar else a[0]
    return buf.arr


def sample_(x, model, scheme, t0, x0, nsteps, **kw):
    """sample!(x, model, scheme, t0, x0, nsteps)   (simulation.jl:62-67): fill the preallocated x."""
    return sample(model, scheme, t0, x0, nsteps, x.shape[0], out=x, **kw)


def _times(t0, dt, nsteps):
    return t0 + dt * np.arange(nsteps + 1)  # t0 .+ scheme.Δt * (0:nsteps)   (simulation.jl:72)


def simulate(model, scheme, t0, x0, nsteps, npaths=None, *, seed=0, precision="f64", math="fast", stream=0,
             path_offset=0, rng=None, out=None, device=False, layout="reference"):
    """simulate(model, scheme, t0, x0, nsteps[, npaths]) -> (x, t)   (simulation.jl:71-76, 104-117, 149-154)

    Full trajectories, start point included.  layout="soa" keeps the coalesced device layout
    (nsteps+1, D, npaths) instead of the reference's."""
    if isinstance(scheme, MultilevelScheme):
        return _ml_simulate(model, scheme, t0, x0, nsteps, npaths, seed=seed, precision=precision, math=math,
                            stream=stream, path_offset=path_offset, rng=rng, device=device)
    L = _lib.lib()
    v0, scalar = _x0(model, x0)
    single = npaths is None
    n = 1 if single else int(npaths)
    D, nrows = model._D, int(nsteps) + 1
    if layout == "reference":
        shape = (n, nrows) if scalar else (n, nrows, D)
        lay = _lib.LAYOUT_REFERENCE
    else:
        shape = (nrows, n) if scalar else (nrows, D, n)
        lay = _lib.LAYOUT_SOA
    buf = _Buf(shape, precision, out, device)
    o = _opts(scheme, t0, seed, precision, math, stream, path_offset, rng=rng)
    p = _params(model)
    _lib.check(L.sdeb200_simulate(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                  int(nsteps), n, lay, buf.ptr))
    x = buf.arr
    if single and layout == "reference":
        x = x[0]  # (nsteps+1,) or (nsteps+1, D): Julia's (nsteps+1,) / (D, nsteps+1)
    return x, _times(t0, scheme.dt, int(nsteps))


def simulate_(x, model, scheme, t0, x0, w=None, *, seed=0, precision=None, math="fast", stream=0, path_offset=0, rng=None):
    """simulate!(x, model, scheme, t0, x0[, w])   (simulation.jl:104-117 and 119-134)

    With `w`: trajectories driven by the cumulative Wiener path w (same layout as x, last axis M); `x is w`
    is allowed, as in the reference.  Without: on-device noise."""
    L = _lib.lib()
    v0, scalar = _x0(model, x0)
    is_torch = hasattr(x, "data_ptr")
    for arr in (x, w):
        if arr is not None and not (arr.is_contiguous() if hasattr(arr, "is_contiguous") else arr.flags.c_contiguous):
            raise ValueError("simulate_ works in place: x and w must be C-contiguous")
    xprec = "f32" if str(x.dtype).endswith("float32") else ("f64" if str(x.dtype).endswith("float64") else None)
    if xprec is None:
        raise TypeError(f"x must be float32 or float64, not {x.dtype}")
    if precision is not None and _PREC[precision] != _PREC[xprec]:
        raise TypeError(f"precision={precision!r} disagrees with x.dtype {x.dtype}: the library would write past the buffer")
    precision = xprec
    if w is not None and str(w.dtype) != str(x.dtype):
        raise TypeError(f"w.dtype {w.dtype} must equal x.dtype {x.dtype} (eltype(w) drives the arithmetic, simulation.jl:119-134)")
    o = _opts(scheme, t0, seed, precision, math, stream, path_offset, rng=rng)
    p = _params(model)
    D, M = model._D, max(model._M, 1)
    xs = tuple(x.shape)
    if len(xs) < (1 if scalar else 2) or (not scalar and xs[-1] != D):
        raise ValueError(f"x {xs} must be ([npaths,] nrows{'' if scalar else ', ' + str(D)})")
    if w is None:
        n = int(np.prod(xs[:-1] if scalar else xs[:-2], dtype=np.int64)) if len(xs) > (1 if scalar else 2) else 1
        nrows = xs[-1] if scalar else xs[-2]
        ptr = x.data_ptr() if is_torch else x.ctypes.data
        _lib.check(L.sdeb200_simulate(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                      nrows - 1, n, _lib.LAYOUT_REFERENCE, ptr))
        return x
    # x: ([npaths,] nrows) for scalar x0, ([npaths,] nrows, D) otherwise; w likewise with M on the last axis
    nrows = xs[-1] if scalar else xs[-2]
    n = int(np.prod(xs[:-1] if scalar else xs[:-2], dtype=np.int64)) if len(xs) > (1 if scalar else 2) else 1
    if int(np.prod(tuple(w.shape), dtype=np.int64)) != n * nrows * M or (not scalar and xs[-1] != D):
        raise ValueError(f"x {xs} and w {tuple(w.shape)} are inconsistent for D={D}, M={M}")
    xp = x.data_ptr() if is_torch else x.ctypes.data
    wp = w.data_ptr() if hasattr(w, "data_ptr") else w.ctypes.data
    _lib.check(L.sdeb200_simulate_wiener(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                         nrows, n, wp, xp))
    return x


def wiener_(w, dt, *, seed=0, stream=0, path_offset=0, rng=None, math="fast"):
    """wiener!(w, Δt)   (simulation.jl:22-35): w[(npaths,) nrows (, dim)], row 0 = 0, running sums of N(0, Δt)."""
    L = _lib.lib()
    is_torch = hasattr(w, "data_ptr")
    if not (w.is_contiguous() if hasattr(w, "is_contiguous") else w.flags.c_contiguous):
        raise ValueError("wiener_ works in place: w must be C-contiguous")
    if not str(w.dtype).endswith(("float32", "float64")):
        raise TypeError(f"w must be float32 or float64, not {w.dtype}")
    precision = "f32" if str(w.dtype).endswith("float32") else "f64"
    o = _lib.Opts()
    o.precision, o.math, o.dt = _PREC[precision], _MATH[math], float(dt)
    o.seed, o.stream, o.path_offset = int(seed), int(stream), int(path_offset)
    o.rng = int(DEFAULT_RNG if rng is None else rng)
    ws = tuple(w.shape)
    if len(ws) == 1:
        n, nrows, dim = 1, ws[0], 1
    elif len(ws) == 2:
        n, nrows, dim = ws[0], ws[1], 1
    else:
        n, nrows, dim = ws
    _lib.check(L.sdeb200_wiener(C.byref(o), dim, nrows, n, w.data_ptr() if is_torch else w.ctypes.data))
    return w


def logtransition(model, scheme, times, data, *, out=None, device=False):
    """logtransition(model, scheme, times, data)   (src/schemes/schemes.jl:64-76 with euler_maruyama.jl:8-13)

    Gaussian Euler–Maruyama log-densities of the transitions data[i-1] -> data[i] along stored trajectories.
    `data`: ([npaths,] nrows) for scalar models, ([npaths,] nrows, D) otherwise (what `simulate` returns);
    `times`: the uniform grid t0 + Δt*(0:nrows-1) (or just t0).  Returns ([npaths,] nrows-1)."""
    L = _lib.lib()
    D = model._D
    is_torch = hasattr(data, "data_ptr")
    precision = "f32" if str(data.dtype).endswith("float32") else "f64"
    shp = tuple(data.shape)
    if D > 1:
        if shp[-1] != D:
            raise ValueError(f"data must carry {D} state components on its last axis")
        shp = shp[:-1]
    single = len(shp) == 1
    n, nrows = (1, shp[0]) if single else (int(np.prod(shp[:-1], dtype=np.int64)), shp[-1])
    t = np.atleast_1d(np.asarray(times, dtype=np.float64))
    if len(t) > 1 and (len(t) != nrows or not np.allclose(np.diff(t), scheme.dt, rtol=1e-9, atol=1e-15)):
        raise ValueError("times must be the uniform grid t0 + scheme.Δt*(0:nrows-1)")
    if not (data.is_contiguous() if is_torch else data.flags.c_contiguous):
        raise ValueError("data must be C-contiguous")
    buf = _Buf((nrows - 1,) if single else (n, nrows - 1), precision, out, device or is_torch)
    o = _opts(scheme, float(t[0]), 0, precision)
    p = _params(model)
    _lib.check(L.sdeb200_logtransition(model._handle, C.byref(o), p.ctypes.data_as(_PD), nrows, n,
                                       data.data_ptr() if is_torch else data.ctypes.data, buf.ptr))
    return buf.arr


# ---------------------------------------------------------------------------------------------------
# fused estimators (new: the reference returns raw states)
# ---------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Moments:
    """Σx_d, Σx_d² of every state component."""
    kind = _lib.PAYOFF_MOMENTS
    comp: int = 0
    strike: float = 0.0
    discount: float = 1.0


@dataclass(frozen=True)
class Call:
    """discount * max(x[comp] - strike, 0)"""
    strike: float
    discount: float = 1.0
    comp: int = 0
    kind = _lib.PAYOFF_CALL


@dataclass(frozen=True)
class Put:
    strike: float
    discount: float = 1.0
    comp: int = 0
    kind = _lib.PAYOFF_PUT


@dataclass(frozen=True)
class Component:
    comp: int = 0
    strike: float = 0.0
    discount: float = 1.0
    kind = _lib.PAYOFF_COMPONENT


def _payoff(p):
    s = _lib.Payoff()
    s.kind, s.comp, s.strike, s.discount = p.kind, int(p.comp), float(p.strike), float(p.discount)
    return s


@dataclass
class Estimate:
    mean: np.ndarray
    stderr: np.ndarray
    n: int
    nonfinite: int
    sums: np.ndarray  # [nf, 2]: Σf, Σf² (exactly accumulated, correctly rounded)

    @staticmethod
    def from_sums(sums, n, bad):
        sums = np.asarray(sums, dtype=np.float64).reshape(-1, 2)
        good = max(n - bad, 1)
        mean = sums[:, 0] / good
        var = np.maximum(sums[:, 1] / good - mean * mean, 0.0) * (good / max(good - 1, 1))
        return Estimate(mean, np.sqrt(var / good), int(n), int(bad), sums)


def estimate(model, scheme, payoff, t0, x0, nsteps, npaths, *, seed=0, precision="f64", math="fast", stream=0,
             path_offset=0, rng=None, terminal=None):
    """Monte Carlo estimate of E[payoff(X_T)]: `sample` with the payoff reduced on the fly (warp shuffles, CTA
    partials, exact accumulation, one NCCL all-reduce when a communicator is up).  `npaths` is this rank's
    share; the returned sums cover all ranks.  Oracle: mean(payoff.(sample(model, scheme, ...)))."""
    if isinstance(scheme, MultilevelScheme):
        from .multilevel import ml_estimate
        return ml_estimate(model, scheme, payoff, t0, x0, nsteps, npaths, seed=seed, precision=precision, math=math,
                           stream=stream, path_offset=path_offset, rng=rng)
    L = _lib.lib()
    v0, _ = _x0(model, x0)
    o = _opts(scheme, t0, seed, precision, math, stream, path_offset, rng=rng)
    p = _params(model)
    nf = model._D if payoff.kind == _lib.PAYOFF_MOMENTS else 1
    res = np.zeros(2 * nf + 2)
    pay = _payoff(payoff)
    tptr = None if terminal is None else terminal.data_ptr()
    _lib.check(L.sdeb200_estimate(model._handle, C.byref(o), p.ctypes.data_as(_PD), v0.ctypes.data_as(_PD),
                                  int(nsteps), int(npaths), C.byref(pay), tptr, res.ctypes.data_as(_PD)))
    return Estimate.from_sums(res[:2 * nf], res[2 * nf], res[2 * nf + 1])
