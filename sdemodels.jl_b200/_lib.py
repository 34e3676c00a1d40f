"""ctypes binding of libsdeb200.so — the C ABI of include/sdeb200.h.

There is deliberately no fallback: if the CUDA library is missing or no device is present the
product fails loudly (a silent CPU path would void every parity and performance claim).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDEB200_LIB") or os.path.join(_HERE, "libsdeb200.so")  # env override: tuning builds only

OK, E_ARG, E_PARSE, E_NVRTC, E_CUDA, E_NCCL, E_DOMAIN, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7
EULER_MARUYAMA, MILSTEIN, RUNGE_KUTTA, EXACT = 0, 1, 2, 3
F64, F32 = 0, 1
MATH_FAST, MATH_STRICT = 0, 1
LAYOUT_REFERENCE, LAYOUT_SOA = 0, 1
RNG_DEFAULT, RNG_V1, RNG_V2 = 0, 1, 2
PAYOFF_MOMENTS, PAYOFF_CALL, PAYOFF_PUT, PAYOFF_COMPONENT = 0, 1, 2, 3
NBINS = 68
SHARD_ALIGN = 1024


class SDEB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sdeb200 error {code}: {msg}")
        self.code = code


class DomainError(SDEB200Error, ArithmeticError):
    """Some path ended non-finite — Julia raises DomainError from sqrt of a negative number."""


class Opts(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("precision", C.c_int32), ("math", C.c_int32), ("rng", C.c_int32),
                ("t0", C.c_double), ("dt", C.c_double), ("seed", C.c_uint64), ("stream", C.c_uint32),
                ("reserved2", C.c_uint32), ("path_offset", C.c_int64)]


class Payoff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("comp", C.c_int32), ("strike", C.c_double), ("discount", C.c_double)]


# name -> (restype, argtypes): every symbol include/sdeb200.h declares
_VP, _I, _I64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
_PD, _PO, _PP = C.POINTER(C.c_double), C.POINTER(Opts), C.POINTER(Payoff)
SYMBOLS = {
    "sdeb200_version": (_I, []),
    "sdeb200_last_error": (C.c_char_p, []),
    "sdeb200_device_count": (_I, [C.POINTER(C.c_int)]),
    "sdeb200_init": (_I, [_I]),
    "sdeb200_init_devices": (_I, [_I, C.POINTER(C.c_int)]),
    "sdeb200_device_info": (_I, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdeb200_shutdown": (_I, []),
    "sdeb200_set_stream": (_I, [_VP]),
    "sdeb200_synchronize": (_I, []),
    "sdeb200_last_kernel_ms": (_D, []),
    "sdeb200_launch_count": (_I64, []),
    "sdeb200_set_tma": (_I, [_I]),
    "sdeb200_tma_launch_count": (_I64, []),
    "sdeb200_set_sample1_max": (_I64, [_I64]),
    "sdeb200_measure_peak": (_I, [_I, _PD, _PD]),
    "sdeb200_model_create": (_I, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), _I, C.POINTER(_VP)]),
    "sdeb200_model_builtin": (_I, [C.c_char_p, C.POINTER(_VP)]),
    "sdeb200_model_free": (None, [_VP]),
    "sdeb200_model_info": (_I, [_VP] + [C.POINTER(C.c_int)] * 4),
    "sdeb200_model_name": (C.c_char_p, [_VP]),
    "sdeb200_model_param_name": (C.c_char_p, [_VP, _I]),
    "sdeb200_model_variable_name": (C.c_char_p, [_VP, _I]),
    "sdeb200_model_drift_expr": (C.c_char_p, [_VP, _I]),
    "sdeb200_model_diffusion_expr": (C.c_char_p, [_VP, _I, _I]),
    "sdeb200_model_cuda_source": (C.c_char_p, [_VP]),
    "sdeb200_model_coefficients": (_I, [_VP, _PD, _D, _PD, _PD, _PD]),
    "sdeb200_model_compile": (_I, [_VP, _I, _I]),
    "sdeb200_sample": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _VP]),
    "sdeb200_simulate": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _I, _VP]),
    "sdeb200_simulate_wiener": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _VP, _VP]),
    "sdeb200_wiener": (_I, [_PO, _I, _I64, _I64, _VP]),
    "sdeb200_logtransition": (_I, [_VP, _PO, _PD, _I64, _I64, _VP, _VP]),
    "sdeb200_mlmc_sample_level": (_I, [_VP, _PO, _PD, _PD, _D, _I64, _I64, _I64, _VP, _VP]),
    "sdeb200_mlmc_simulate_level": (_I, [_VP, _PO, _PD, _PD, _D, _I64, _I64, _I64, _VP, _VP]),
    "sdeb200_estimate": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _PP, _VP, _PD]),
    "sdeb200_mlmc_estimate_level": (_I, [_VP, _PO, _PD, _PD, _D, _I64, _I64, _I64, _PP, _PD]),
    "sdeb200_mlmc_estimate": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _I, _I64, C.POINTER(C.c_int64), _PP, _PD]),
    "sdeb200_buffer_alloc": (_I, [_I, _I, C.POINTER(C.c_int64), C.POINTER(_VP)]),
    "sdeb200_buffer_ptr": (_VP, [_VP]),
    "sdeb200_buffer_info": (_I, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sdeb200_buffer_copy_to_host": (_I, [_VP, _I64, _I64, _VP]),
    "sdeb200_buffer_copy_from_host": (_I, [_VP, _I64, _I64, _VP]),
    "sdeb200_buffer_free": (_I, [_VP]),
    "sdeb200_mlmc_estimate_sharded": (_I, [_VP, _PO, _PD, _PD, _I64, _I64, _I, _I64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _PP, _PD]),
    "sdeb200_comm_unique_id": (_I, [_VP]),
    "sdeb200_comm_init": (_I, [_VP, _I, _I]),
    "sdeb200_comm_destroy": (_I, []),
    "sdeb200_comm_info": (_I, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdeb200_bins_to_double": (_D, [C.POINTER(C.c_int64)]),
    "sdeb200_bins_add": (None, [C.POINTER(C.c_int64), _D]),
}

_lib = None


def lib():
    """Load libsdeb200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                              "(sdeb200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = lib().sdeb200_last_error().decode("utf-8", "replace")
    if rc == E_DOMAIN:
        raise DomainError(rc, msg)
    raise SDEB200Error(rc, msg)
