"""FP32 kernels against an FP32 oracle, and the FP32-vs-FP64 deviation as a measured table.

The reference is Float64 only; the FP32 kernels are an addition of the new framework.  Round 1 compared them
with the Float64 oracle under loosened tolerances (2e-4 at the end of 1024 steps), which cannot tell a kernel error from
the conditioning of FP32 arithmetic.  Here the two are separated:

 (1) oracle/sde_oracle.c is compiled a second time with `real = float` (oracle_f32_*): the SAME loops, every operation a
     float operation, -ffp-contract=off.  The STRICT FP32 kernels must equal it BIT FOR BIT on Wiener-driven
     trajectories — built-in functors, NVRTC-generated functors, EM / Milstein / RungeKutta, ragged shapes, in place.
 (2) The deviation of FP32 trajectories from the Float64 reference arithmetic ON IDENTICAL INCREMENTS (a float-valued
     Wiener path: its differences are exact in both precisions) is measured per row index and written to
     gpurun_out/fp32_deviation.json (kept under profiles/); the assertions are the measured envelope, see DESIGN.md §3.
"""
import json
import os

import numpy as np
import pytest

from conftest import BS_P, HESTON_P
from test_dsl import CASES, MODELS

pytestmark = pytest.mark.gpu
EM, MIL, RK = 0, 1, 2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


def model(S, name, params):
    if name in ("BlackScholes", "Heston", "OrnsteinUhlenbeck"):
        return getattr(S, name)(*params)
    if name not in _cache:
        _cache[name] = S.sde_model(name + "_g32", MODELS.get(name) or S.REGISTRY[name])
    return _cache[name](*params)


def wiener32(O, npaths, nsteps, dim, dt, seed):
    z = np.random.default_rng(seed).standard_normal((npaths, nsteps, dim)).astype(np.float32)
    return O.wiener_z(dt, z, precision="f32")


@pytest.mark.parametrize("name,params,x,t", [c for c in CASES if c[0] != "Deterministic"])
def test_strict_fp32_kernels_are_bit_identical_to_the_fp32_oracle(gpu, O, name, params, x, t):
    S = gpu
    D, M, _ = O.dims(name)
    m = model(S, name, params)
    for npaths, nsteps in ((257, 129), (33, 40)):
        dt = 1.0 / 128
        w = wiener32(O, npaths, nsteps, M, dt, 5)
        for sid in [EM] + ([MIL, RK] if M == 1 else []):
            ref = O.simulate_wiener(name, sid, params, t, dt, x, w, precision="f32")
            assert ref.dtype == np.float32
            out = np.empty_like(ref)
            xs = out[..., 0] if D == 1 else out
            ws = np.ascontiguousarray(w[..., 0] if (M == 1 and D == 1) else w)
            sch = (S.EulerMaruyama, S.Milstein, S.RungeKutta)[sid](dt)
            S.simulate_(xs, m, sch, t, x[0] if D == 1 else x, ws, math="strict")
            assert np.array_equal(out, ref), (name, sid, npaths)


def test_strict_fp32_in_place_and_device_resident(gpu, O):
    import torch
    S = gpu
    for name, params, x0, D in (("Heston", HESTON_P, [100.0, 0.4], 2), ("BlackScholes", BS_P, 100.0, 1)):
        nsteps, npaths, dt = 513, 300, 1.0 / 512          # long enough for the TMA tile form
        w = wiener32(O, npaths, nsteps, D, dt, 9)
        ref = O.simulate_wiener(name, EM, params, 0.0, dt, x0, w, precision="f32")
        buf = torch.from_numpy(w[..., 0].copy() if D == 1 else w.copy()).cuda()
        c0 = S.tma_launch_count()
        S.simulate_(buf, getattr(S, name)(*params), S.EulerMaruyama(dt), 0.0, x0, buf, math="strict")
        assert S.tma_launch_count() > c0
        assert np.array_equal(buf.cpu().numpy().reshape(ref.shape), ref)


def test_fp32_deviation_table(gpu, O):
    """FP32 vs the Float64 reference arithmetic on IDENTICAL increments, per row index, 10^4 paths x 1024 steps.
    north_star: "on identical pre-generated Wiener increments ... within 1e-5 (FP32)".  With a float-valued Wiener path the
    increments w[i] - w[i-1] are exact in both precisions, and the FP32 trajectories stay within 1e-5 of the Float64 ones
    over all 1024 steps (measured max 6.2e-6 for the float oracle, which (1) pins the STRICT kernels to): the deviation grows
    like sqrt(row) x 2^-24 (independent roundings of the state).  Round 1's 2e-4 came from feeding the FP32 kernel a ROUNDED
    copy of a Float64 path, i.e. from different increments, not from the kernels."""
    S = gpu
    rows = [1, 8, 32, 128, 512, 1024]
    table = {}
    for name, params, x0, sid, D in (("BlackScholes", BS_P, 100.0, EM, 1), ("BlackScholes", BS_P, 100.0, MIL, 1),
                                     ("Heston", HESTON_P, [100.0, 0.4], EM, 2)):
        nsteps, npaths, dt = 1024, 10000, 1.0 / 1024
        w32 = wiener32(O, npaths, nsteps, D, dt, 3)
        ref = O.simulate_wiener(name, sid, params, 0.0, dt, x0, w32.astype(np.float64))          # Float64 arithmetic
        sch = (S.EulerMaruyama, S.Milstein)[sid](dt)
        m = getattr(S, name)(*params)
        for math in ("strict", "fast"):
            out = np.empty((npaths, nsteps + 1, D), dtype=np.float32)
            S.simulate_(out[..., 0] if D == 1 else out, m, sch, 0.0, x0, np.ascontiguousarray(w32[..., 0] if D == 1 else w32), math=math)
            dev = np.abs(out.astype(np.float64) - ref) / np.abs(ref)
            key = f"{name}_{('em', 'mil')[sid]}_{math}"
            table[key] = {str(r): {"max": float(dev[:, r].max()), "p99": float(np.quantile(dev[:, r].max(axis=-1), 0.99)),
                                   "median": float(np.median(dev[:, r].max(axis=-1)))} for r in rows}
            assert dev.max() <= 1e-5, (key, dev.max())                 # north_star's FP32 tolerance, every row
            eps = 2.0 ** -24
            for r in rows:
                assert dev[:, r].max() <= 8 * eps * np.sqrt(r) + 4 * eps, (key, r, dev[:, r].max())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp32_deviation.json"), "w") as fh:
        json.dump({"what": "max / p99 / median over 10^4 paths of |x_fp32 - x_fp64| / |x_fp64| at row r, identical float-valued "
                           "Wiener path, Wiener-driven simulate!", "rows": rows, "table": table}, fh, indent=1)
