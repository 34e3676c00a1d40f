"""Multi-GPU plumbing: one process per GPU (torchrun), paths sharded contiguously, one NCCL communicator
owned by libsdeb200.  torch.distributed is used only to pass the NCCL unique id around and for barriers."""
import ctypes as C
import os

import numpy as np

from . import _lib


def shard_range(npaths, rank, world, align=_lib.SHARD_ALIGN):
    """Contiguous global path range [lo, hi) of `rank`: boundaries on multiples of `align` paths so that the
    chunking of the payoff reduction — and therefore every rounding — is independent of the number of GPUs."""
    units = -(-npaths // align)
    lo = min(npaths, (units * rank // world) * align)
    hi = min(npaths, (units * (rank + 1) // world) * align)
    return lo, hi


def init(device=None):
    """Bind this process to its GPU (LOCAL_RANK) and, under torchrun, build the NCCL communicator."""
    L = _lib.lib()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
    _lib.check(L.sdeb200_init(local))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(L.sdeb200_comm_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _lib.check(L.sdeb200_comm_init(raw, world, rank))
    return rank, world


def init_devices(n=0):
    """One process owning several GPUs (sdeb200_init_devices): bind devices 0..n-1 (n <= 0: all).  Afterwards every call
    on host arrays is sharded over them — contiguous path ranges, one worker thread per GPU, exact sums combined on the
    host — with results identical to a single GPU.  Returns the number of devices bound."""
    out = C.c_int()
    _lib.check(_lib.lib().sdeb200_init_devices(int(n), C.byref(out)))
    return out.value


def device_info():
    n, p = C.c_int(), C.c_int()
    _lib.lib().sdeb200_device_info(C.byref(n), C.byref(p))
    return n.value, p.value


class DeviceBuffer:
    """A library-owned device buffer (sdeb200_buffer_*): results that stay on the GPU for callers without a device array
    type.  `.ptr` goes wherever a driver takes a data pointer (`out=` accepts the object itself)."""

    def __init__(self, shape, precision="f64"):
        self.shape = tuple(int(v) for v in shape)
        self.precision = precision
        self.dtype = np.float64 if precision in ("f64", "float64") else np.float32
        arr = (C.c_int64 * len(self.shape))(*self.shape)
        h = C.c_void_p()
        _lib.check(_lib.lib().sdeb200_buffer_alloc(_lib.F64 if self.dtype == np.float64 else _lib.F32, len(self.shape), arr, C.byref(h)))
        self._h = h
        self.ptr = _lib.lib().sdeb200_buffer_ptr(h)

    def data_ptr(self):
        return self.ptr

    def to_host(self, offset=0, count=None):
        n = int(np.prod(self.shape, dtype=np.int64)) - offset if count is None else int(count)
        out = np.empty(n, dtype=self.dtype)
        _lib.check(_lib.lib().sdeb200_buffer_copy_to_host(self._h, int(offset), n, out.ctypes.data))
        return out.reshape(self.shape) if (offset == 0 and count is None) else out

    def from_host(self, arr, offset=0):
        a = np.ascontiguousarray(arr, dtype=self.dtype)
        _lib.check(_lib.lib().sdeb200_buffer_copy_from_host(self._h, int(offset), a.size, a.ctypes.data))

    def free(self):
        if self._h:
            _lib.lib().sdeb200_buffer_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def comm_info():
    n, r = C.c_int(), C.c_int()
    _lib.lib().sdeb200_comm_info(C.byref(n), C.byref(r))
    return n.value, r.value


def combine_bins_host(per_rank_bins):
    """Host-side twin of the all-reduce: add int64 limb arrays of several ranks and round once.
    Used by the gloo (CPU) tests of the multi-rank logic."""
    total = np.sum(np.asarray(per_rank_bins, dtype=np.int64), axis=0)
    L = _lib.lib()
    return L.sdeb200_bins_to_double(total.ctypes.data_as(C.POINTER(C.c_int64)))
