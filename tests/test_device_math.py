"""The device header's math (Philox block function, fused normal transforms) compiled with g++ and checked on the
CPU against the independent statement in oracle/philox_spec.c.  Runs without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "host_math.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "host_math_harness.cpp")], check=True)
    return C.CDLL(so)


def test_philox_matches_random123_vectors(H, O):
    for ctr, key in (([0] * 4, [0] * 2), ([0xffffffff] * 4, [0xffffffff] * 2),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
                     ([1, 2, 3, 4], [5, 6])):
        out = (C.c_uint32 * 4)()
        H.h_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert list(out) == O.philox4x32_10(ctr, key)


def test_fp64_normal_transform_matches_the_specification(H, O):
    n = 300000
    out = np.zeros(n)
    H.h_normals_f64(C.c_uint64(5), C.c_uint32(1), C.c_uint64(77), C.c_uint64(0), C.c_int64(n), C.c_double(1.0),
                    out.ctypes.data_as(C.c_void_p))
    ref = O.normals(5, 1, 77, 0, n)
    err = np.abs(out - ref)
    # |z| error is bounded by the conditioning of sqrt(-2 ln u) near u -> 1 (tiny radii): 1e-16 / |z|
    assert np.all(err <= 2e-16 * np.maximum(1.0, 1.0 / np.maximum(np.abs(ref), 1e-3)) + 5e-16 * np.abs(ref))
    assert err.max() < 1e-13
    # scaled draw: N(0, dt)
    dt = 1.0 / 1024
    H.h_normals_f64(C.c_uint64(5), C.c_uint32(1), C.c_uint64(77), C.c_uint64(0), C.c_int64(n), C.c_double(dt),
                    out.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(out - ref * np.sqrt(dt))) < 1e-14


def test_fp64_transform_edge_bits(H):
    """tails (many leading zeros), u1 -> 1 (clamped radius), exact quadrant boundaries: always finite, and
    the pair lies on the circle of the right radius."""
    cases = [(0, 0, 0, 0), (1, 0, 0, 0), (0xffffffff, 0xffffffff, 0, 0), (0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff),
             (5, 0, 0, 0x40000000), (5, 0, 0, 0x80000000), (5, 0, 0, 0xC0000000), (0, 1, 0xffffffff, 0x1fffffff),
             (0, 1, 0, 0x20000000), (123, 0x80000000, 0xdeadbeef, 0x9abcdef0)]
    for r in cases:
        z = (C.c_double * 2)()
        H.h_pair_from_bits((C.c_uint32 * 4)(*r), C.c_double(1.0), z)
        assert np.isfinite(z[0]) and np.isfinite(z[1])
        X1 = (r[1] << 32) | r[0]
        w = 2.0 * 1087 * np.log(2.0) if X1 == 0 else -2.0 * (np.log(float(X1)) - 64 * np.log(2.0))
        assert abs(np.hypot(z[0], z[1]) - np.sqrt(w + 2.0 ** -50)) <= 1e-13 * max(1.0, np.sqrt(w)) + 2e-8 * (w < 1e-12)
    # angle check at a quadrant: X2 = 2^62 is a quarter turn -> (cos, sin) = (0, 1)
    z = (C.c_double * 2)()
    H.h_pair_from_bits((C.c_uint32 * 4)(0, 0x40000000, 0, 0x40000000), C.c_double(1.0), z)
    assert abs(z[0]) < 1e-15 and z[1] > 0


def test_fp32_normal_transform_matches_the_specification(H, O):
    n = 200000
    out = np.zeros(n, dtype=np.float32)
    H.h_normals_f32(C.c_uint64(9), C.c_uint32(2), C.c_uint64(4), C.c_uint64(0), C.c_int64(n), C.c_float(1.0),
                    out.ctypes.data_as(C.c_void_p))
    ref = O.normals(9, 2, 4, 0, n, "f32")
    assert np.max(np.abs(out - ref)) < 2e-5
    assert np.all(np.isfinite(out))


def test_tma_tile_geometry(H):
    """sde_tma_group / sde_tma_head (sde_args.h): g consecutive paths make the tensor map's row stride a multiple of the
    alignment — 16 bytes (what TMA requires; K3t, K7t, K2u with 128-byte tiles) or 32 bytes (K2u with 64-byte tiles: whole
    DRAM sectors) — and the box of sub-path j starts at the first aligned time row."""
    H.h_tma_group.argtypes = [C.c_int64, C.c_int]
    H.h_tma_head.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int]
    for align in (16, 32):
        for es in (4, 8):
            for D in (1, 2, 4):
                rowb = D * es
                for nrows in list(range(1, 70)) + [252, 253, 512, 513, 1024, 1025, 4097]:
                    pitch = nrows * rowb
                    g = H.h_tma_group(pitch, align)
                    assert g in (1, 2, 4, 8)[:3 if align == 16 else 4] and (g * pitch) % align == 0
                    assert g == 1 or ((g // 2) * pitch) % align != 0
                    heads = [H.h_tma_head(j, pitch, rowb, align) for j in range(g)]
                    assert heads[0] == 0
                    for j, h in enumerate(heads):
                        assert (j * pitch + h * rowb) % align == 0                       # aligned box start
                        assert all((j * pitch + k * rowb) % align != 0 for k in range(h))  # and the first such row
                        assert h <= max(0, align // rowb - 1)                               # HMAX of the kernels


def test_per_path_philox_form_is_bit_identical(H, O):
    """philox4x32_10(PhiloxPath, blk): rounds 0 and 1 partly hoisted per path — same 128 bits as the plain block function."""
    rng = np.random.default_rng(3)
    H.h_philox_path.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    for _ in range(200):
        path = int(rng.integers(0, 2 ** 63)) if rng.random() < 0.5 else int(rng.integers(0, 2 ** 20))
        stream, blk = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))
        key = [int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))]
        out = (C.c_uint32 * 4)()
        H.h_philox_path(path, stream, blk, (C.c_uint32 * 2)(*key), out)
        assert list(out) == O.philox4x32_10([path & 0xffffffff, path >> 32, blk, stream], key)


def test_log_fast_and_rcp_fast(H):
    """The log-density kernel's logarithm (table + degree-4 polynomial of the normal generator) and reciprocal (seed + two
    Newton steps) against libm in extended precision: absolute error of the logarithm <= 4e-16 + 1 ulp, reciprocal <= 1 ulp;
    zero, subnormal, negative, infinite and NaN arguments take the library routine."""
    rng = np.random.default_rng(11)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), 1.0 + rng.uniform(-1e-3, 1e-3, 50000),
                        rng.uniform(0.5, 2.0, 100000), [1.0, 2.0, 0.5, 2.2250738585072014e-308, 1.7976931348623157e308,
                                                        1.0 - 2.0 ** -53, 1.0 + 2.0 ** -52]])
    out = np.zeros_like(x)
    H.h_log_fast(x.ctypes.data_as(C.c_void_p), C.c_int64(x.size), out.ctypes.data_as(C.c_void_p))
    ref = np.log(x.astype(np.longdouble))
    err = np.abs(out.astype(np.longdouble) - ref).astype(np.float64)
    assert np.all(err <= 4e-16 + np.spacing(np.abs(ref).astype(np.float64))), float(err.max())
    special = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310])
    so = np.zeros_like(special)
    H.h_log_fast(special.ctypes.data_as(C.c_void_p), C.c_int64(special.size), so.ctypes.data_as(C.c_void_p))
    with np.errstate(all="ignore"):
        want = np.log(special)
    assert np.array_equal(so, want, equal_nan=True)
    y = np.concatenate([x, -x[:1000], [1e-300, 1e300, 3.0, -7.0]])
    ro = np.zeros_like(y)
    H.h_rcp_fast(y.ctypes.data_as(C.c_void_p), C.c_int64(y.size), ro.ctypes.data_as(C.c_void_p))
    rref = (np.longdouble(1.0) / y.astype(np.longdouble))
    rerr = np.abs(ro.astype(np.longdouble) - rref).astype(np.float64)
    assert np.all(rerr <= np.spacing(np.abs(rref).astype(np.float64))), float((rerr / np.abs(rref).astype(np.float64)).max())
    sp = np.array([0.0, np.inf, -np.inf, 5e-324, 1e-309, 1.7e308])
    spo = np.zeros_like(sp)
    H.h_rcp_fast(sp.ctypes.data_as(C.c_void_p), C.c_int64(sp.size), spo.ctypes.data_as(C.c_void_p))
    with np.errstate(all="ignore"):
        assert np.array_equal(spo, 1.0 / sp)


@pytest.mark.parametrize("version", [1, 2])
def test_heston_polar_step_equals_the_plain_step(H, version):
    """Heston's polar step (FAST math, FP64: ONE square root of V*w instead of sqrt(V) and sqrt(w)) against the plain step on
    the same pair, on both normal streams: at most one ulp of the new state apart, equal at V = 0, non-finite for V < 0."""
    rng = np.random.default_rng(12)
    params = (C.c_double * 5)(0.01, 10.0, 0.3, 0.1, 0.4)
    dt, worst = 1.0 / 1024, 0.0
    xp, xq = (C.c_double * 2)(), (C.c_double * 2)()
    for i in range(4000):
        w4 = (C.c_uint32 * 4)(*[int(v) for v in rng.integers(0, 2 ** 32, 4)])
        S, V = 100.0 * np.exp(rng.standard_normal()), (0.4 * rng.random() ** 4 if i % 50 else 0.0)
        H.h_heston_polar(version, params, C.c_double(dt), w4, (C.c_double * 2)(S, V), xp, xq)
        assert xp[0] != -1.0
        for d in (0, 1):
            worst = max(worst, abs(xp[d] - xq[d]) / (2.0 ** -52 * abs(xq[d])) if xq[d] != 0 else abs(xp[d]))
        if V == 0.0:
            assert xp[0] == xq[0] and xp[1] == xq[1]
    assert worst <= 1.5, worst   # measured: 1.0 ulp of the new state
    H.h_heston_polar(version, params, C.c_double(dt), (C.c_uint32 * 4)(1, 2, 3, 4), (C.c_double * 2)(100.0, -0.01), xp, xq)
    assert not np.isfinite(xp[0]) and not np.isfinite(xp[1])


def test_heston_polar_step_fp32(H):
    """The FP32 polar step against the plain FP32 step on the same pair (host twin of the FP32 stream): within two float
    ulps of the new state, equal at V = 0, non-finite for V < 0."""
    rng = np.random.default_rng(13)
    params = (C.c_float * 5)(0.01, 10.0, 0.3, 0.1, 0.4)
    dt, worst = 1.0 / 1024, 0.0
    xp, xq = (C.c_float * 2)(), (C.c_float * 2)()
    for i in range(4000):
        w2 = (C.c_uint32 * 2)(*[int(v) for v in rng.integers(0, 2 ** 32, 2)])
        S, V = 100.0 * np.exp(rng.standard_normal()), (0.4 * rng.random() ** 4 if i % 50 else 0.0)
        H.h_heston_polar_f32(params, C.c_float(dt), w2, (C.c_float * 2)(S, V), xp, xq)
        assert xp[0] != -1.0
        for d in (0, 1):
            worst = max(worst, abs(xp[d] - xq[d]) / (2.0 ** -23 * abs(xq[d])) if xq[d] != 0 else abs(xp[d]))
        if V == 0.0:
            assert xp[0] == xq[0] and xp[1] == xq[1]
    assert worst <= 2.0, worst
    H.h_heston_polar_f32(params, C.c_float(dt), (C.c_uint32 * 2)(1, 2), (C.c_float * 2)(100.0, -0.01), xp, xq)
    assert not np.isfinite(xp[0]) and not np.isfinite(xp[1])


def test_heston_coarse_step_from_two_polar_triples(H):
    """Heston::step_polar2 (the coarse step of a coupled level with two substeps taken straight from the two polar triples
    of its fine steps) against the reference form — the increments summed from zero, then one plain step: a few ulps of
    the new state apart, equal at V = 0, non-finite for V < 0."""
    rng = np.random.default_rng(14)
    params = (C.c_double * 5)(0.01, 10.0, 0.3, 0.1, 0.4)
    dtf, worst = 1.0 / 1024, 0.0
    xp, xq = (C.c_double * 2)(), (C.c_double * 2)()
    for i in range(4000):
        wa = (C.c_uint32 * 3)(*[int(v) for v in rng.integers(0, 2 ** 32, 3)])
        wb = (C.c_uint32 * 3)(*[int(v) for v in rng.integers(0, 2 ** 32, 3)])
        S, V = 100.0 * np.exp(rng.standard_normal()), (0.4 * rng.random() ** 4 if i % 50 else 0.0)
        H.h_heston_polar2(params, C.c_double(dtf), wa, wb, (C.c_double * 2)(S, V), xp, xq)
        for d in (0, 1):
            worst = max(worst, abs(xp[d] - xq[d]) / (2.0 ** -52 * abs(xq[d])) if xq[d] != 0 else abs(xp[d]))
        if V == 0.0:
            assert xp[0] == xq[0] and xp[1] == xq[1]
    assert worst <= 4.0, worst
    H.h_heston_polar2(params, C.c_double(dtf), (C.c_uint32 * 3)(1, 2, 3), (C.c_uint32 * 3)(4, 5, 6), (C.c_double * 2)(100.0, -0.01), xp, xq)
    assert not np.isfinite(xp[0]) and not np.isfinite(xp[1])


def test_fp32_mean_reversion_keeps_its_level(H):
    """FAST math, FP32: the variance row reverts as V + κΔt (θ - V), not as V (1 - κΔt) + κθΔt — with 1 - κΔt rounded to
    24 bits the second form carries κΔt to only 2^-24 / (κΔt) relative and moves the level the state reverts to.  The
    level θ is an exact fixed point of the FP32 drift for any κ, Δt, and a step from 2θ is the once-rounded 2θ - κΔt θ;
    Float64 uses the one-FMA form, where the same effect stays at rounding level."""
    H.h_heston_drift_f32.restype = C.c_float
    H.h_heston_drift_f64.restype = C.c_double
    rng = np.random.default_rng(15)
    f32 = np.float32
    for _ in range(300):
        kappa, theta = float(f32(rng.uniform(0.1, 12.0))), float(f32(rng.uniform(0.01, 0.5)))
        dt = float(f32(2.0 ** -rng.integers(6, 17) * rng.uniform(0.5, 1.0)))
        p32 = (C.c_float * 5)(0.01, kappa, theta, 0.1, 0.4)
        assert H.h_heston_drift_f32(p32, C.c_float(dt), C.c_float(theta)) == f32(theta)
        v1 = float(H.h_heston_drift_f32(p32, C.c_float(dt), C.c_float(2 * theta)))
        kdt = float(f32(f32(kappa) * f32(dt)))
        exact = 2 * theta - kdt * theta                       # in double, from the float-rounded κΔt
        assert abs(v1 - exact) <= 2.0 ** -23 * 2 * theta      # one float rounding of the result (the expanded form: up to 2^-24 θ / 1 per unit of κΔt more)
        p64 = (C.c_double * 5)(0.01, kappa, theta, 0.1, 0.4)
        assert abs(H.h_heston_drift_f64(p64, C.c_double(dt), C.c_double(theta)) / theta - 1.0) <= 3e-16
