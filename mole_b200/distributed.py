"""Host plumbing for walker sharding across ranks.

Walkers shard in contiguous ranges; the only data-path exchange is a fp64 sum of the 62-double packed
accumulator vector.  On GPUs this is `mole_acc_allreduce` (NCCL inside libmole_b200.so); the helpers here
serve callers that bring their own collective (torch.distributed, any backend) and the CPU/gloo tests.
"""
import ctypes as C

import numpy as np

from . import ffi


def shard(n_walkers_global, world_size, rank):
    """Contiguous walker range of `rank`: (n_local, walker_offset).  Global ids key the Philox stream, so
    results do not depend on world_size."""
    base, rem = divmod(int(n_walkers_global), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return n_local, offset


def acc_to_array(acc):
    """mole_acc_host -> packed float64[62] (the device layout, ACC_* in mole_internal.h)."""
    return np.frombuffer(bytes(acc), dtype=np.float64, count=ffi.AccHost.N_DOUBLES).copy()


def array_to_acc(arr, n_params):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    assert arr.size == ffi.AccHost.N_DOUBLES
    acc = ffi.AccHost()
    C.memmove(C.byref(acc), arr.ctypes.data, arr.nbytes)
    acc.n_params = int(n_params)
    acc.reserved = 0
    return acc


def allreduce_acc(acc, group=None):
    """Sum the reduced moments over all ranks with torch.distributed (concatenate_worker_data,
    src/vmc/src/vmc.rs:108-130, as a 496-byte allreduce)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(acc_to_array(acc))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return array_to_acc(t.cpu().numpy(), acc.n_params)
