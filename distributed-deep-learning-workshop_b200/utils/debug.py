"""Debug aids: cross-rank checksums to prove replicas are bit-identical after an all-reduce."""
from __future__ import annotations

import torch


def checksum_across_ranks(t: torch.Tensor) -> bool:
    """True when every rank holds exactly the same bytes in `t` (sum + xor-folded int view compared across ranks)."""
    from ..parallel import core

    v = t.detach().contiguous().view(-1)
    iv = v.view(torch.int32) if v.element_size() == 4 else v.view(torch.int16).to(torch.int32)
    sig = torch.stack([iv.sum(dtype=torch.int64), (iv.to(torch.int64) * 2654435761 % 4294967291).sum()]).to(torch.float64)
    if core.size() == 1:
        return True
    lo = sig.clone()
    hi = sig.clone()
    import torch.distributed as dist

    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))
