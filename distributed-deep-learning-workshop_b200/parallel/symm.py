"""Symmetric (peer-mapped + NVLS-multicast) device memory for the fused all-reduce kernels.

Allocation, handle exchange between the rank processes and multicast binding (cuMulticast*) are delegated to
`torch.distributed._symmetric_memory`; what we take from it is only raw addresses: the device array of every
rank's buffer pointer and the multicast pointer.  Kernels in csrc/allreduce.cu do all the data movement.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

_flag_cache: Dict[int, "SymmetricBuffer"] = {}
_live = []


def reset() -> None:
    _flag_cache.clear()
    _bcast_stage.clear()
    _live.clear()


def end_job() -> None:
    """Persistent rank processes (Runner(persistent=True)): drop the data buffers of the job that just finished - every
    comm kernel ends with a cross-rank barrier, so no peer still touches them - but keep the process group's flag
    buffers, which the next job reuses."""
    keep = {id(b) for b in _flag_cache.values()} | {id(v[0]) for v in _bcast_stage.values()}
    _live[:] = [b for b in _live if id(b) in keep]


class SymmetricBuffer:
    """A tensor allocated at the same virtual offset on every rank, mapped into every rank's address space."""

    def __init__(self, numel: int, dtype: torch.dtype, device: Optional[torch.device] = None, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tensor = symm_mem.empty(numel, dtype=dtype, device=dev)
        self.tensor.zero_()
        try:
            self.handle = symm_mem.rendezvous(self.tensor, group=self.group)
        except TypeError:
            self.handle = symm_mem.rendezvous(self.tensor, self.group.group_name)
        h = self.handle
        self.rank = int(h.rank)
        self.world = int(h.world_size)
        self.peer_ptrs_dev = int(h.buffer_ptrs_dev)
        self.mc_ptr = int(getattr(h, "multicast_ptr", 0) or 0)
        self.peer_ptrs = [int(p) for p in h.buffer_ptrs]
        torch.cuda.synchronize()
        dist.barrier(group=self.group, device_ids=[torch.cuda.current_device()])
        _live.append(self)

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0


def flag_buffer(group=None) -> SymmetricBuffer:
    """One zero-initialised symmetric flag array per process group (barrier slots of the comm kernels)."""
    from .. import ops

    g = group if group is not None else dist.group.WORLD
    key = id(g)
    if key not in _flag_cache:
        words = int(ops.ext("_b200_comm").FLAG_WORDS)
        _flag_cache[key] = SymmetricBuffer(words, torch.int32, group=g)
    return _flag_cache[key]


def make_comm(buf: SymmetricBuffer, group=None):
    """csrc Comm object bound to a symmetric data buffer + the group's flag buffer."""
    from .. import ops

    ext = ops.ext("_b200_comm")
    flags = flag_buffer(group)
    return ext.Comm(buf.peer_ptrs_dev, flags.peer_ptrs_dev, buf.mc_ptr, buf.rank, buf.world)


_bcast_stage: Dict[int, tuple] = {}


def broadcast_tensor(t: torch.Tensor, root: int = 0, group=None, stage_mb: int = 64, blocks: int = 16) -> torch.Tensor:
    """In-place broadcast of a CUDA tensor with OUR multicast / P2P-store kernel (csrc/allreduce.cu broadcast_kernel,
    SURVEY.md K2 - the reference's `BroadcastGlobalVariablesCallback(0)`, P1/03:305-308): the tensor is streamed through
    a cached symmetric staging buffer in `stage_mb` chunks; the kernel's entry / exit flag barriers order the copies."""
    g = group if group is not None else dist.group.WORLD
    key = id(g)
    if key not in _bcast_stage:
        buf = SymmetricBuffer(stage_mb * 2 ** 20 // 4, torch.float32, t.device, group=g)
        _bcast_stage[key] = (buf, make_comm(buf, g))
    buf, comm = _bcast_stage[key]
    flat = t.detach().reshape(-1)
    if not flat.is_contiguous():
        raise ValueError("broadcast_tensor needs a contiguous tensor")
    raw = flat.view(torch.uint8)
    stage = buf.tensor.view(torch.uint8)
    cap = stage.numel()
    for lo in range(0, raw.numel(), cap):
        n = min(cap, raw.numel() - lo)
        n16 = (n + 15) // 16 * 16
        if buf.rank == root:
            stage[:n].copy_(raw[lo:lo + n])
        comm.broadcast(0, n16 // 4, "f32", root, blocks)
        if buf.rank != root:
            raw[lo:lo + n].copy_(stage[:n])
    return t
