"""Device plumbing: torch owns device memory, streams and (optionally) torch.distributed.

Nothing here computes on the hot path; it moves numpy arrays to HBM once and hands raw
pointers + the current CUDA stream to libslb200.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _native


def device():
    _native.require_device()
    if not torch.cuda.is_available():
        raise _native.NativeLibraryError("torch sees no CUDA device; safe_learning_b200 has no "
                                         "CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(array, dtype=torch.float64):
    """numpy (or tensor) -> contiguous device tensor of `dtype`."""
    if isinstance(array, torch.Tensor):
        return array.to(device=device(), dtype=dtype).contiguous()
    np_dtype = {torch.float64: np.float64, torch.int64: np.int64, torch.uint8: np.uint8,
                torch.int32: np.int32}[dtype]
    host = np.ascontiguousarray(np.asarray(array), dtype=np_dtype)
    return torch.from_numpy(host).to(device())


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=device())


def stream():
    """The current CUDA stream as the void* libslb200 expects."""
    return torch.cuda.current_stream().cuda_stream


def ptr(tensor):
    return None if tensor is None else tensor.data_ptr()


# ---- torch.distributed (one process per GPU; NCCL on GPUs, gloo in the CPU tests) ----------
def dist_info():
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(nindex, rank=None, world=None):
    """Contiguous flat-index range [begin, end) of this rank."""
    if rank is None or world is None:
        rank, world = dist_info()
    per = -(-nindex // world)
    begin = min(rank * per, nindex)
    return begin, min(begin + per, nindex)


def allgather_rows(row):
    """All-gather one small 1-D tensor per rank -> [world, len] (device of `row`); a view of the
    input when torch.distributed is not initialised.  The only collective of a Lyapunov sweep."""
    rank, world = dist_info()
    if world == 1:
        return row.view(1, -1)
    import torch.distributed as dist
    out = torch.empty((world, row.numel()), dtype=row.dtype, device=row.device)
    dist.all_gather_into_tensor(out, row.view(1, -1).contiguous())
    return out


# ---- peer-memory key exchange (slb_exchange): symmetric buffers over NVLink --------------------
_EXCHANGE = {"tried": False, "struct": None, "keep": None, "error": None}


def get_exchange():
    """``slb_exchange`` of this process (one per process, shared by every sharded sweep), or None
    when the ranks cannot map each other's memory (no NVLink peer access, CPU/gloo runs, or
    ``SLB200_EXCHANGE=nccl``): callers then fall back to an NCCL all-gather of the keys.

    Collective on first use: every rank allocates ``slb_fail_key[2][world]`` in symmetric memory
    (``torch.distributed._symmetric_memory``: CUDA VMM allocations exchanged between the processes
    of the NVSwitch domain), zeroes it, and all ranks rendezvous."""
    import os
    if _EXCHANGE["tried"]:
        return _EXCHANGE["struct"]
    _EXCHANGE["tried"] = True
    rank, world = dist_info()
    if world == 1 or world > _native.SLB_MAX_RANKS or not torch.cuda.is_available():
        return None
    if os.environ.get("SLB200_EXCHANGE", "peer") != "peer":
        return None
    import torch.distributed as dist
    ok = 1
    handle = buf = None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(2 * world * 4, dtype=torch.int64, device=device())
        buf.zero_()
        handle = symm_mem.rendezvous(buf, dist.group.WORLD)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        if len(ptrs) != world or any(p == 0 for p in ptrs):
            ok = 0
    except Exception as exc:  # pragma: no cover - depends on the machine
        _EXCHANGE["error"] = repr(exc)
        ok = 0
    # every rank must take the same path: agree (NCCL, once per process)
    flag = torch.tensor([ok], dtype=torch.int32, device=device())
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None
    seq = torch.zeros(1, dtype=torch.int64, device=device())
    x = _native.SlbExchange()
    x.world, x.rank = world, rank
    for r, p in enumerate(ptrs):
        x.slots[r] = p
    x.seq_dev = seq.data_ptr()
    torch.cuda.synchronize()
    dist.barrier()                      # every rank's slots are zero before anyone pushes
    _EXCHANGE["struct"] = x
    _EXCHANGE["keep"] = (buf, handle, seq)
    return x


_FACTOR_DEPENDENCY = [0]


def note_factor_dependency():
    """A restore registered an event the packed-factor launches wait for (functions.PackedCache):
    CUDA graphs of sweeps captured before that have no wait node and must be re-captured."""
    if _FACTOR_DEPENDENCY[0] == 0:
        _FACTOR_DEPENDENCY[0] = 1


def factor_dependency_epoch():
    return _FACTOR_DEPENDENCY[0]
