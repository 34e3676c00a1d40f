"""`@sde_model` and the model registry (src/models/codegen.jl, src/models/models.jl).

The macro's work — splitting the equations into drift and diffusion, folding the Cholesky factor of the
Wiener correlation matrix into the diffusion, ordering the parameters — happens in the C++ front end behind
`sdeb200_model_create`, which also emits the CUDA functor compiled by NVRTC.  This module only mirrors the
Julia surface: `sde_model(...)` returns a model *type* whose instances carry the parameter values.
"""
import ctypes as C

import numpy as np

from . import _lib


class AbstractSDE:
    """Instances are what Julia's generated `struct Name <: AbstractSDE{D,M}` instances are."""
    _handle = None
    _name = ""
    _fields = ()
    _D = _M = 0
    _kind = 0

    def __init__(self, *args, **kw):
        vals = list(args)
        if kw:
            vals += [kw[f] for f in self._fields[len(vals):]]
        if len(vals) != len(self._fields):
            raise TypeError(f"{self._name} takes {len(self._fields)} parameters {self._fields}, got {len(vals)}")
        self.params = np.array([float(v) for v in vals], dtype=np.float64)
        for f, v in zip(self._fields, self.params):
            object.__setattr__(self, f, float(v))

    def __repr__(self):
        return f"{self._name}({', '.join(repr(float(v)) for v in self.params)})"


def _finish(cls, handle):
    L = _lib.lib()
    D, M, NP, K = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(L.sdeb200_model_info(handle, C.byref(D), C.byref(M), C.byref(NP), C.byref(K)))
    cls._handle = handle
    cls._D, cls._M, cls._kind = D.value, M.value, K.value
    cls._fields = tuple(L.sdeb200_model_param_name(handle, i).decode() for i in range(NP.value))
    cls._vars = tuple(L.sdeb200_model_variable_name(handle, i).decode() for i in range(D.value))
    return cls


def sde_model(name, equations, *parameter_vars):
    """@sde_model Name <equations> [parameters...]   (codegen.jl:157-159)

    `equations`: the macro body as text, one equation per line, e.g.
        sde_model("Heston", '''dS = r*S*dt + sqrt(V)*S*dW1
                               dV = κ*(θ-V)*dt + σ*sqrt(V)*dW2
                               dW1*dW2 = ρ*dt''')
    """
    L = _lib.lib()
    h = C.c_void_p()
    arr = (C.c_char_p * max(len(parameter_vars), 1))(*[p.encode() for p in parameter_vars])
    _lib.check(L.sdeb200_model_create(name.encode(), equations.encode(), arr, len(parameter_vars), C.byref(h)))
    cls = type(name, (AbstractSDE,), {"_name": name, "_equations": equations})
    return _finish(cls, h)


def _builtin(name):
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.sdeb200_model_builtin(name.encode(), C.byref(h)))
    cls = type(name, (AbstractSDE,), {"_name": name})
    return _finish(cls, h)


# registry equations, verbatim from src/models/black_scholes.jl:1, models.jl:39-54,
# cox_ingersoll_ross.jl:1, ornstein_uhlenbeck.jl:1
REGISTRY = {
    "BlackScholes": "dS = r*S*dt + σ*S*dW",
    "Heston": "dS = r*S*dt + sqrt(V)*S*dW1\ndV = κ*(θ-V)*dt + σ*sqrt(V)*dW2\ndW1*dW2 = ρ*dt",
    "CoxIngersollRoss": "dr = κ*(θ-r)*dt + σ*sqrt(r)*dW",
    "OrnsteinUhlenbeck": "dx = θ*(μ-x)*dt + σ*dW",
    "FitzHughNagumo": "dX = ɛ*(X-X^3-Y+s)*dt + σ1*dW1\ndY = (γ*X-Y+β)*dt + σ2*dW2",
    "Lorenz": "dx = σ*(y-x)*dt + σ1*dw1\ndy = (x*(ρ-z)-y)*dt + σ2*dw2\ndz = (x*y-β*z)*dt + σ3*dw3",
}
_cache = {}


def registry_model(name, generated=False):
    """Model types of the reference's registry.  BlackScholes, Heston and OrnsteinUhlenbeck default to the
    hand-written ahead-of-time functors (the first and the last also carry their Exact samplers);
    `generated=True` builds them from their equations like any other model."""
    key = (name, generated)
    if key not in _cache:
        if name in ("BlackScholes", "Heston", "OrnsteinUhlenbeck") and not generated:
            _cache[key] = _builtin(name)
        else:
            _cache[key] = sde_model(name, REGISTRY[name])
    return _cache[key]


def dim(model):
    """dim(model) == (D, M)   (models.jl:3-5)"""
    return (model._D, model._M)


def variables(model):
    """variables(model): the state symbols in order (codegen.jl:149)"""
    return model._vars


def fieldnames(model_type):
    return tuple(model_type._fields)


def _coefficients(model, t, x):
    L = _lib.lib()
    D, M = model._D, model._M
    xx = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))
    if xx.shape != (D,):
        raise ValueError(f"state must have {D} components")
    mu = np.zeros(D)
    sig = np.zeros((D, max(M, 1)))
    p = model.params if len(model.params) else np.zeros(1)
    PD = C.POINTER(C.c_double)
    _lib.check(L.sdeb200_model_coefficients(model._handle, p.ctypes.data_as(PD), float(t), xx.ctypes.data_as(PD),
                                            mu.ctypes.data_as(PD), sig.ctypes.data_as(PD)))
    return mu, sig


def drift(model, t, x):
    """drift(model, t, x): scalar for D == 1, else a D-vector (codegen.jl:68-74, 185-189) — evaluated by the
    compiled device functor."""
    mu, _ = _coefficients(model, t, x)
    return float(mu[0]) if model._D == 1 else mu


def diffusion(model, t, x):
    """diffusion(model, t, x): scalar for a 1x1 matrix, else D x M with the correlation Cholesky folded in
    (codegen.jl:105-112); `0` for models without Wiener processes."""
    _, sig = _coefficients(model, t, x)
    if model._M == 0:
        return 0.0
    if sig.size == 1:
        return float(sig[0, 0])
    return sig
