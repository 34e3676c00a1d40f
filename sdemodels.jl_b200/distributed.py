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
