"""The C-ABI library loads without a GPU, exports every symbol include/sdeb200.h declares, agrees with the
header's constants, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sdeb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdeb200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(S):
    L = S._lib.lib()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sdeb200.h but not exported"
        assert n in S._lib.SYMBOLS, f"{n} has no ctypes signature in _lib.py"
    assert set(S._lib.SYMBOLS) <= set(names)


def test_header_constants_match_python(S):
    txt = open(os.path.join(ROOT, "include", "sdeb200.h")).read()
    consts = dict(re.findall(r"#define\s+(SDEB200_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", txt))
    L = S._lib
    assert int(consts["SDEB200_VERSION"]) == L.lib().sdeb200_version()
    assert int(consts["SDEB200_NBINS"]) == L.NBINS and int(consts["SDEB200_SHARD_ALIGN"]) == L.SHARD_ALIGN
    for k, v in (("E_ARG", L.E_ARG), ("E_PARSE", L.E_PARSE), ("E_NVRTC", L.E_NVRTC), ("E_CUDA", L.E_CUDA),
                 ("E_NCCL", L.E_NCCL), ("E_DOMAIN", L.E_DOMAIN), ("E_UNSUPPORTED", L.E_UNSUPPORTED),
                 ("MILSTEIN", L.MILSTEIN), ("RUNGE_KUTTA", L.RUNGE_KUTTA), ("EXACT", L.EXACT), ("F32", L.F32), ("MATH_STRICT", L.MATH_STRICT), ("LAYOUT_SOA", L.LAYOUT_SOA),
                 ("PAYOFF_CALL", L.PAYOFF_CALL), ("PAYOFF_PUT", L.PAYOFF_PUT), ("PAYOFF_COMPONENT", L.PAYOFF_COMPONENT)):
        assert int(consts["SDEB200_" + k]) == v


def test_struct_layouts(S):
    assert C.sizeof(S._lib.Opts) == 56 and C.sizeof(S._lib.Payoff) == 24


def _has_gpu(S):
    n = C.c_int()
    return S._lib.lib().sdeb200_device_count(C.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback(S):
    if _has_gpu(S):
        pytest.skip("a GPU is present")
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(S.BlackScholes(0.01, 0.3), S.EulerMaruyama(0.01), 0.0, 100.0, 10, 5)
    assert ei.value.code == S._lib.E_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(S.SDEB200Error):
        S.estimate(S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.EulerMaruyama(0.01), S.Moments(), 0.0, [100.0, 0.4], 4, 8)


def test_argument_validation_needs_no_device(S):
    m = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4)
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(m, S.Milstein(0.01), 0.0, [100.0, 0.4], 10, 10)
    assert ei.value.code == S._lib.E_UNSUPPORTED and "milstein.jl:4" in str(ei.value)
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(m, S.EulerMaruyama(-1.0), 0.0, [100.0, 0.4], 10, 10)
    assert ei.value.code == S._lib.E_ARG
    with pytest.raises(ValueError):
        S.sample(m, S.EulerMaruyama(0.1), 0.0, 100.0, 10, 10)  # x0 must have D components


def test_exact_accumulator_host_half(S):
    """bins_add / bins_to_double: the exact sum, correctly rounded, independent of the order of additions."""
    import math
    L = S._lib.lib()
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(5000) * 1e8, rng.standard_normal(5000) * 1e-8, [1e300, -1e300, 5e-324, 2.5e-310]])
    P64 = C.POINTER(C.c_int64)
    for perm in (vals, vals[::-1], rng.permutation(vals)):
        bins = np.zeros(S._lib.NBINS, dtype=np.int64)
        for v in perm:
            L.sdeb200_bins_add(bins.ctypes.data_as(P64), float(v))
        assert L.sdeb200_bins_to_double(bins.ctypes.data_as(P64)) == math.fsum(vals)
    bins = np.zeros(S._lib.NBINS, dtype=np.int64)
    assert L.sdeb200_bins_to_double(bins.ctypes.data_as(P64)) == 0.0
    for v in (1.0, -1.0, 2.0 ** 60, -3.5e-300, 1.7976931348623157e308, 0.1):
        b = np.zeros(S._lib.NBINS, dtype=np.int64)
        L.sdeb200_bins_add(b.ctypes.data_as(P64), v)
        assert L.sdeb200_bins_to_double(b.ctypes.data_as(P64)) == v
    # ties-to-even: 2^53 + 1 is not representable; 2^53 + 1 + 1 is
    b = np.zeros(S._lib.NBINS, dtype=np.int64)
    for v in (2.0 ** 53, 1.0):
        L.sdeb200_bins_add(b.ctypes.data_as(P64), v)
    assert L.sdeb200_bins_to_double(b.ctypes.data_as(P64)) == 2.0 ** 53
    L.sdeb200_bins_add(b.ctypes.data_as(P64), 1.0)
    assert L.sdeb200_bins_to_double(b.ctypes.data_as(P64)) == 2.0 ** 53 + 2


def test_milstein_without_a_symbolic_derivative_is_refused(S):
    """when the front end cannot differentiate the diffusion (max, min, sign, ...), Milstein(Δt) must fail with
    E_UNSUPPORTED instead of silently running Euler–Maruyama; EulerMaruyama and RungeKutta stay available.  The refusal
    happens before any device is touched, so this runs on the CPU."""
    import numpy as np
    cls = S.sde_model("Floored", "dS = r*S*dt + σ*max(S, 0.5)*dW")
    m = cls(0.01, 0.3)
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(m, S.Milstein(0.01), 0.0, 1.0, 4, 10)
    assert ei.value.code == S._lib.E_UNSUPPORTED and "max" in str(ei.value) and "RungeKutta" in str(ei.value)
    ok = S.sde_model("Smooth", "dS = r*S*dt + σ*sqrt(S)*dW")(0.01, 0.3)
    for sch in (S.Milstein(0.01), S.RungeKutta(0.01), S.EulerMaruyama(0.01)):
        with pytest.raises(S.SDEB200Error) as ei:       # no GPU here: the call gets past the scheme check and fails on the device
            S.sample(ok, sch, 0.0, 1.0, 4, 10)
        assert ei.value.code in (S._lib.E_CUDA, S._lib.E_NVRTC)
    with pytest.raises(S.SDEB200Error) as ei:
        S.sample(m, S.RungeKutta(0.01), 0.0, 1.0, 4, 10)
    assert ei.value.code in (S._lib.E_CUDA, S._lib.E_NVRTC)
