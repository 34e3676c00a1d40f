"""sdemodels.jl_b200 — B200-native Monte Carlo path simulation behind the SDEModels.jl API.

Only the reference's data-parallel hot path lives here: `sample` / `simulate` over the Euler–Maruyama and
Milstein steps for `@sde_model` models, and the coupled fine/coarse sampler of MultilevelScheme.  The names
mirror the reference (Julia's `f!` is spelled `f_`); the compute is CUDA behind the C ABI in
include/sdeb200.h (csrc/).  Import as `sdemodels_jl_b200` via the repo-root loader `_pkg.py`.
"""
from . import _lib
from ._lib import DomainError, SDEB200Error
from .distributed import DeviceBuffer, comm_info, device_info, init as init_distributed, init_devices, shard_range
from .models import (AbstractSDE, REGISTRY, diffusion, dim, drift, fieldnames, registry_model, sde_model,
                     variables)
from .multilevel import (AdaptiveMLMC, MultilevelEstimate, MultilevelScheme, cost, level_stream, ml_estimate, mlmc_adaptive,
                         multilevel_sample_level, multilevel_simulate_level, npaths, schemes)
from .schemes import EulerMaruyama, Exact, Milstein, RungeKutta, subdivide
from .simulation import (Call, Component, Estimate, Moments, Put, estimate, logtransition, sample, sample_, simulate,
                         simulate_, wiener_)


def __getattr__(name):
    # registry model types (BlackScholes, Heston, CoxIngersollRoss, ...) are created on first use
    if name in REGISTRY:
        return registry_model(name)
    raise AttributeError(name)


def measure_peak(precision="f64"):
    """FMA-chain micro-benchmark on the bound GPU -> (TFLOP/s, ms)."""
    import ctypes as C
    t, ms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().sdeb200_measure_peak(0 if precision == "f64" else 1, C.byref(t), C.byref(ms)))
    return t.value, ms.value


def last_kernel_ms():
    return _lib.lib().sdeb200_last_kernel_ms()


def launch_count():
    return _lib.lib().sdeb200_launch_count()


def set_tma(enabled):
    """Tuning switch: False routes trajectory kernels with a TMA (tensor-tile) form to the plain tile kernels."""
    return bool(_lib.lib().sdeb200_set_tma(1 if enabled else 0))


def set_sample1_max(npaths):
    """Tuning switch: terminal-value launches of at most `npaths` paths use the one-path-per-thread kernel form."""
    return _lib.lib().sdeb200_set_sample1_max(int(npaths))


def tma_launch_count():
    return _lib.lib().sdeb200_tma_launch_count()


def set_stream(cuda_stream_ptr):
    _lib.check(_lib.lib().sdeb200_set_stream(cuda_stream_ptr))


def synchronize():
    _lib.check(_lib.lib().sdeb200_synchronize())
