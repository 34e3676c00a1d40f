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
    """sde_tma_group / sde_tma_head (sde_args.h): g consecutive paths make the tensor map's row stride a multiple of
    16 bytes (TMA requirement), and the box of sub-path j starts at the first 16-byte-aligned time row."""
    H.h_tma_group.argtypes = [C.c_int64]
    H.h_tma_head.argtypes = [C.c_int, C.c_int64, C.c_int]
    for es in (4, 8):
        for D in (1, 2, 4):
            rowb = D * es
            for nrows in list(range(1, 70)) + [252, 253, 512, 513, 1024, 1025, 4097]:
                pitch = nrows * rowb
                g = H.h_tma_group(pitch)
                assert g in (1, 2, 4) and (g * pitch) % 16 == 0 and (g == 1 or ((g // 2) * pitch) % 16 != 0)
                heads = [H.h_tma_head(j, pitch, rowb) for j in range(g)]
                assert heads[0] == 0
                for j, h in enumerate(heads):
                    assert (j * pitch + h * rowb) % 16 == 0                       # aligned box start
                    assert all((j * pitch + k * rowb) % 16 != 0 for k in range(h))  # and the first such row
                    assert h <= (0 if rowb % 16 == 0 else (1 if rowb % 8 == 0 else 3))  # HMAX of the kernels


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
