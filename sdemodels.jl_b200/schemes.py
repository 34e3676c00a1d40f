"""Scheme types of the accelerated path (src/schemes/schemes.jl:3-11,34-37,51-54)."""
from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class _UnconditionalScheme:
    """`struct T <: UnconditionalScheme; Δt::Float64 end` with the `T(Δt, t, s) = T(Δt)` constructor (schemes.jl:3-11)."""
    dt: float

    def __init__(self, dt, t=None, s=None):
        object.__setattr__(self, "dt", float(dt))

    @property
    def Δt(self):
        return self.dt


class EulerMaruyama(_UnconditionalScheme):
    """src/schemes/euler_maruyama.jl:1-6"""
    _id = _lib.EULER_MARUYAMA


class Milstein(_UnconditionalScheme):
    """src/schemes/milstein.jl:4-10 — models with one Wiener process only, as in the reference."""
    _id = _lib.MILSTEIN


class RungeKutta(_UnconditionalScheme):
    """src/schemes/runge_kutta.jl:4-11 — derivative-free Milstein (diffusion re-evaluated at a predictor state);
    models with one Wiener process only, as in the reference."""
    _id = _lib.RUNGE_KUTTA


class Exact(_UnconditionalScheme):
    """Closed-form one-step sampling (src/simulation.jl:50-57,78-89): BlackScholes (black_scholes.jl:3-13) and
    OrnsteinUhlenbeck (ornstein_uhlenbeck.jl:3-13) — the built-in model types."""
    _id = _lib.EXACT


def subdivide(scheme, nsubsteps):
    """subdivide(scheme::T, n) = T(scheme.Δt / n)   (schemes.jl:53-54)"""
    return type(scheme)(scheme.dt / nsubsteps)
