"""ctypes binding of the CPU oracle (oracle/sde_oracle.c, oracle/philox_spec.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Array conventions follow Julia's memory order: a Julia
``Array{SVector{D,Float64}}(nrows, npaths)`` is the numpy C-order array ``[npaths, nrows, D]``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsde_oracle.so")

EULER_MARUYAMA, MILSTEIN, RUNGE_KUTTA, EXACT = 0, 1, 2, 3


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("sde_oracle.c", "philox_spec.c", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "libsde_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_model_name.restype = C.c_char_p
        L.oracle_level_dt.restype = C.c_double
        L.oracle_level_dt.argtypes = [C.c_double, C.c_int64, C.c_int64]
        L.oracle_npaths.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_double, _ip]
        L.oracle_npaths.restype = None
        L.oracle_cost.argtypes = [C.c_int64, C.c_int64, C.c_int64, _ip, _ip]
        L.oracle_cost.restype = None
        L.oracle_normals_f64.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int64, _dp]
        L.oracle_normals_f32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int64, _dp]
        L.oracle_normals_f64_v2.argtypes = L.oracle_normals_f64.argtypes
        L.oracle_normals_f64.restype = L.oracle_normals_f32.restype = L.oracle_normals_f64_v2.restype = None
        _lib = L
    return _lib


def model_id(name):
    L = lib()
    for i in range(L.oracle_model_count()):
        if L.oracle_model_name(i).decode() == name:
            return i
    raise KeyError(name)


def dims(model):
    D, M, P = C.c_int(), C.c_int(), C.c_int()
    lib().oracle_model_dims(model_id(model), C.byref(D), C.byref(M), C.byref(P))
    return D.value, M.value, P.value


def _vec(a):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))


def _pv(a):
    a = _vec(a) if a is not None and len(np.atleast_1d(a)) else np.zeros(1)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def drift(model, params, t, x):
    D, M, _ = dims(model)
    out = np.zeros(D)
    p, xx = _pv(params), _vec(x)
    lib().oracle_drift(model_id(model), _ptr(p), C.c_double(t), _ptr(xx), _ptr(out))
    return out


def diffusion(model, params, t, x):
    D, M, _ = dims(model)
    out = np.zeros((D, max(M, 1)))
    p, xx = _pv(params), _vec(x)
    lib().oracle_diffusion(model_id(model), _ptr(p), C.c_double(t), _ptr(xx), _ptr(out))
    return out


def _rt(precision):
    """(numpy dtype, symbol prefix) of the arithmetic the loops run in: the reference's Float64, or float — the same
    loops restated in FP32 (oracle_f32_*), the bit-exact model of the STRICT FP32 kernels."""
    return (np.float32, "oracle_f32_") if precision in ("f32", "float32") else (np.float64, "oracle_")


def step(model, scheme, params, dt, t, x0, dw, precision="f64"):
    """step(model, scheme, t0, x0, Δw): src/schemes/euler_maruyama.jl:1-6, milstein.jl:4-10, runge_kutta.jl:4-11."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    p, xx = _pv(params), _vec(x0)
    dww = np.ascontiguousarray(np.atleast_1d(np.asarray(dw, dtype=rt)))
    out = np.zeros(D, dtype=rt)
    rc = getattr(lib(), pre + "step")(model_id(model), scheme, _ptr(p), C.c_double(dt), C.c_double(t), _ptr(xx), _ptr(dww), _ptr(out))
    if rc:
        raise RuntimeError(f"oracle_step rc={rc}")
    return out


def sample_z(model, scheme, params, t0, dt, x0, z, precision="f64"):
    """simulation.jl:39-48,59-67 with supplied normals z[npaths, nsteps, max(M,1)] -> [npaths, D]."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    z = np.ascontiguousarray(z, dtype=rt)
    npaths, nsteps = z.shape[0], z.shape[1]
    out = np.zeros((npaths, D), dtype=rt)
    p, xx = _pv(params), _vec(x0)
    rc = getattr(lib(), pre + "sample_z")(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt), _ptr(xx),
                               C.c_int64(nsteps), C.c_int64(npaths), _ptr(z), _ptr(out))
    if rc:
        raise RuntimeError(f"oracle_sample_z rc={rc}")
    return out


def wiener_z(dt, z, precision="f64"):
    """wiener! (simulation.jl:22-35): z[npaths, nrows-1, dim] -> w[npaths, nrows, dim]."""
    rt, pre = _rt(precision)
    z = np.ascontiguousarray(z, dtype=rt)
    npaths, nm1, dim = z.shape
    w = np.zeros((npaths, nm1 + 1, dim), dtype=rt)
    getattr(lib(), pre + "wiener_z")(C.c_double(dt), C.c_int64(nm1 + 1), C.c_int64(npaths), C.c_int(dim), _ptr(z), _ptr(w))
    return w


def simulate_z(model, scheme, params, t0, dt, x0, z, precision="f64"):
    """simulate! (simulation.jl:104-117): z[npaths, nsteps, max(M,1)] -> x[npaths, nsteps+1, D]."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    z = np.ascontiguousarray(z, dtype=rt)
    npaths, nsteps = z.shape[0], z.shape[1]
    x = np.zeros((npaths, nsteps + 1, D), dtype=rt)
    p, xx = _pv(params), _vec(x0)
    rc = getattr(lib(), pre + "simulate_z")(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt), _ptr(xx),
                                 C.c_int64(nsteps + 1), C.c_int64(npaths), _ptr(z), _ptr(x))
    if rc:
        raise RuntimeError(f"oracle_simulate_z rc={rc}")
    return x


def simulate_wiener(model, scheme, params, t0, dt, x0, w, inplace=False, precision="f64"):
    """simulate!(x, model, scheme, t0, x0, w) (simulation.jl:119-134): w[npaths, nrows, M] -> x[npaths, nrows, D]."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    w = np.ascontiguousarray(w, dtype=rt)
    npaths, nrows = w.shape[0], w.shape[1]
    x = w if inplace else np.zeros((npaths, nrows, D), dtype=rt)
    p, xx = _pv(params), _vec(x0)
    rc = getattr(lib(), pre + "simulate_wiener")(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt), _ptr(xx),
                                      C.c_int64(nrows), C.c_int64(npaths), _ptr(w), _ptr(x))
    if rc:
        raise RuntimeError(f"oracle_simulate_wiener rc={rc}")
    return x


def logtransition(model, params, t0, dt, x):
    """logtransition(model, EulerMaruyama(dt), times, data) (schemes.jl:64-76): x[npaths, nrows, D] -> [npaths, nrows-1]."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    npaths, nrows = x.shape[0], x.shape[1]
    out = np.zeros((npaths, nrows - 1))
    p = _pv(params)
    rc = lib().oracle_logtransition(model_id(model), _ptr(p), C.c_double(t0), C.c_double(dt), C.c_int64(nrows),
                                    C.c_int64(npaths), _ptr(x), _ptr(out))
    if rc:
        raise RuntimeError(f"oracle_logtransition rc={rc}")
    return out


def level_dt(base_dt, substeps, level):
    return lib().oracle_level_dt(base_dt, substeps, level)


def npaths(substeps, level_lo, level_hi, budget):
    out = np.zeros(level_hi - level_lo + 1, dtype=np.int64)
    lib().oracle_npaths(substeps, level_lo, level_hi, float(budget), out)
    return out


def cost(substeps, level_lo, level_hi, n):
    n = np.ascontiguousarray(n, dtype=np.int64)
    out = np.zeros_like(n)
    lib().oracle_cost(substeps, level_lo, level_hi, n, out)
    return out


def multilevel_sample_z(model, scheme, params, t0, dt_coarse, dt_fine, substeps, x0, z, precision="f64"):
    """_multilevel_sample (multilevel.jl:65-89): z[npaths, ncoarse*substeps, M] -> (xc, xf) each [npaths, D]."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    z = np.ascontiguousarray(z, dtype=rt)
    npaths, nfine = z.shape[0], z.shape[1]
    assert nfine % substeps == 0
    xc, xf = np.zeros((npaths, D), dtype=rt), np.zeros((npaths, D), dtype=rt)
    p, xx = _pv(params), _vec(x0)
    rc = getattr(lib(), pre + "multilevel_sample_z")(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt_coarse),
                                          C.c_double(dt_fine), C.c_int64(substeps), _ptr(xx),
                                          C.c_int64(nfine // substeps), C.c_int64(npaths), _ptr(z), _ptr(xc), _ptr(xf))
    if rc:
        raise RuntimeError(f"oracle_multilevel_sample_z rc={rc}")
    return xc, xf


def multilevel_simulate_z(model, scheme, params, t0, dt_coarse, dt_fine, substeps, x0, z, precision="f64"):
    """_multilevel_simulate (multilevel.jl:91-100): z[npaths, ncoarse*substeps, D] -> (xc, xf) full paths."""
    D, M, _ = dims(model)
    rt, pre = _rt(precision)
    z = np.ascontiguousarray(z, dtype=rt)
    npaths, nfine = z.shape[0], z.shape[1]
    nsteps = nfine // substeps
    xc, xf = np.zeros((npaths, nsteps + 1, D), dtype=rt), np.zeros((npaths, nfine + 1, D), dtype=rt)
    p, xx = _pv(params), _vec(x0)
    rc = getattr(lib(), pre + "multilevel_simulate_z")(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt_coarse),
                                            C.c_double(dt_fine), C.c_int64(substeps), _ptr(xx), C.c_int64(nsteps),
                                            C.c_int64(npaths), _ptr(z), _ptr(xc), _ptr(xf))
    if rc:
        raise RuntimeError(f"oracle_multilevel_simulate_z rc={rc}")
    return xc, xf


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def normals(seed, stream, path, q0, n, precision="f64", rng=1):
    """The sdeb200 normal stream: normals q0..q0+n-1 of one path's sub-stream.  rng = 1 | 2 selects the
    stream version (FP64 only; the FP32 stream is the same in both)."""
    out = np.zeros(n)
    L = lib()
    fn = (L.oracle_normals_f64_v2 if rng == 2 else L.oracle_normals_f64) if precision == "f64" else L.oracle_normals_f32
    fn(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(path), C.c_uint64(q0), C.c_int64(n), out)
    return out


def normals_paths(seed, stream, path0, npaths, nsteps, M, precision="f64", rng=1):
    """z[npaths, nsteps, M] as the RNG-driven kernels consume them (flat index q = step*M + j)."""
    z = np.zeros((npaths, nsteps, max(M, 1)))
    if M > 0:
        for k in range(npaths):
            z[k] = normals(seed, stream, path0 + k, 0, nsteps * M, precision, rng).reshape(nsteps, M)
    return z


def max_threads():
    return lib().oracle_max_threads()


def sample_rng(model, scheme, params, t0, dt, x0, nsteps, npaths_, seed=0, threads=0, want_out=False):
    """CPU baseline: scalar-per-path loops with a ziggurat generator.  Returns (sums[2D], threads, out|None)."""
    D, M, _ = dims(model)
    sums = np.zeros(2 * D)
    out = np.zeros((npaths_, D)) if want_out else None
    p, xx = _pv(params), _vec(x0)
    used = lib().oracle_sample_rng(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt), _ptr(xx),
                                   C.c_int64(nsteps), C.c_int64(npaths_), C.c_uint64(seed), C.c_int(threads),
                                   _ptr(sums), _ptr(out) if want_out else None)
    return sums, used, out


def simulate_rng(model, scheme, params, t0, dt, x0, nsteps, npaths_, seed=0, threads=0):
    D, M, _ = dims(model)
    x = np.zeros((npaths_, nsteps + 1, D))
    p, xx = _pv(params), _vec(x0)
    used = lib().oracle_simulate_rng(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt), _ptr(xx),
                                     C.c_int64(nsteps + 1), C.c_int64(npaths_), C.c_uint64(seed), C.c_int(threads),
                                     _ptr(x))
    return x, used


def multilevel_rng(model, scheme, params, t0, dt_coarse, dt_fine, substeps, x0, ncoarse, npaths_, seed=0, threads=0):
    D, M, _ = dims(model)
    xc, xf = np.zeros((npaths_, D)), np.zeros((npaths_, D))
    p, xx = _pv(params), _vec(x0)
    used = lib().oracle_multilevel_rng(model_id(model), scheme, _ptr(p), C.c_double(t0), C.c_double(dt_coarse),
                                       C.c_double(dt_fine), C.c_int64(substeps), _ptr(xx), C.c_int64(ncoarse),
                                       C.c_int64(npaths_), C.c_uint64(seed), C.c_int(threads), _ptr(xc), _ptr(xf))
    return xc, xf, used
