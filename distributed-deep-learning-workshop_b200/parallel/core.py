"""Process-group bootstrap and the small off-hot-path collectives.

`init()` reads the torchrun-style environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT) that
`Runner` (or torchrun itself) sets.  Without it, the job is a single local process: `size() == 1` and every
collective is the identity - the reference's "test the distributed code on the driver only" rung
(`HorovodRunner(np=-1)`, P1/03:385-395).  GPU jobs use the NCCL backend, CPU jobs gloo (README:3 promises both).
"""
from __future__ import annotations

import datetime
import os
from typing import Any, List, Optional

import torch
import torch.distributed as dist

_state = {"initialized": False, "rank": 0, "size": 1, "local_rank": 0, "backend": None, "owns_pg": False}


def init(backend: Optional[str] = None, timeout_s: float = 600.0) -> None:
    """hvd.init() (reference P1/03:283)."""
    if _state["initialized"]:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    lrk = int(os.environ.get("LOCAL_RANK", str(rk)))
    use_cuda = torch.cuda.is_available() and os.environ.get("B200DDL_FORCE_CPU", "0") != "1"
    if use_cuda:
        # pin this process to its GPU (reference: set_visible_devices(gpus[hvd.local_rank()]), P1/03:291-295)
        torch.cuda.set_device(lrk % torch.cuda.device_count())
    if world > 1:
        if backend is None:
            backend = "nccl" if use_cuda else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
            dist.init_process_group(backend, rank=rk, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
            _state["owns_pg"] = True
    _state.update(initialized=True, rank=rk, size=world, local_rank=lrk, backend=backend if world > 1 else None)


def shutdown() -> None:
    from . import symm

    symm.reset()
    if _state["owns_pg"] and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    _state.update(initialized=False, rank=0, size=1, local_rank=0, backend=None, owns_pg=False)


def is_initialized() -> bool:
    return _state["initialized"]


def rank() -> int:
    return _state["rank"]


def size() -> int:
    return _state["size"]


def local_rank() -> int:
    return _state["local_rank"]


def backend() -> Optional[str]:
    return _state["backend"]


def device() -> torch.device:
    if torch.cuda.is_available() and os.environ.get("B200DDL_FORCE_CPU", "0") != "1":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def barrier() -> None:
    if size() > 1:
        if _state["backend"] == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def allreduce(t: torch.Tensor, average: bool = True) -> torch.Tensor:
    """Out-of-band all-reduce for small tensors (metrics, K3); returns a new tensor."""
    out = t.clone()
    if size() > 1:
        dist.all_reduce(out)
        if average:
            out = out / size()
    return out


def broadcast(t: torch.Tensor, root: int = 0) -> torch.Tensor:
    """In-place broadcast (initial weights / optimizer state / BN statistics, K2).  CUDA tensors on the NCCL backend go
    through our multicast broadcast kernel over symmetric memory (`symm.broadcast_tensor`); everything else (CPU / gloo,
    non-contiguous or tiny tensors, `B200DDL_BROADCAST=nccl`) uses the library collective."""
    if size() > 1:
        use_kernel = (t.is_cuda and _state["backend"] == "nccl" and t.is_contiguous() and t.numel() * t.element_size() >= 1024
                      and os.environ.get("B200DDL_BROADCAST", "sym") != "nccl")
        if use_kernel:
            try:
                from . import symm

                symm.broadcast_tensor(t, root)
                return t
            except (RuntimeError, AttributeError, ImportError) as ex:  # no symmetric memory on this platform
                if not _state.get("warned_bcast"):
                    _state["warned_bcast"] = True
                    print(f"[b200ddl] symmetric broadcast unavailable ({type(ex).__name__}: {str(ex)[:120]}); using NCCL",
                          flush=True)
        dist.broadcast(t, root)
    return t


def broadcast_object(obj: Any, root: int = 0) -> Any:
    if size() == 1:
        return obj
    box = [obj if rank() == root else None]
    dist.broadcast_object_list(box, src=root, device=device() if _state["backend"] == "nccl" else None)
    return box[0]


def allgather_object(obj: Any) -> List[Any]:
    if size() == 1:
        return [obj]
    out: List[Any] = [None] * size()
    dist.all_gather_object(out, obj)
    return out
