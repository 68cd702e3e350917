"""Flatten the parameters of any `torch.nn.Module` into one fp32 buffer (+ one gradient buffer) so that the fused
flat optimizers and the bucketed `DistributedOptimizer` work for arbitrary autograd models (e.g. the reference's
MobileNetV2 + Dense head) exactly as they do for the hand-scheduled ResNet-50 engine."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class FlatParams:
    def __init__(self, module: torch.nn.Module, grads: Optional[torch.Tensor] = None):
        self.module = module
        # gradients become ready roughly in reverse registration order -> lay the buffer out that way
        self.named = [(n, p) for n, p in reversed(list(module.named_parameters())) if p.requires_grad]
        self.ranges: List[Tuple[int, int]] = []
        off = 0
        for _, p in self.named:
            self.ranges.append((off, off + _align(p.numel())))
            off += _align(p.numel())
        self.numel = max(off, 64)
        dev = self.named[0][1].device if self.named else torch.device("cpu")
        self.params = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.grads = grads if grads is not None else torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self._hooks = []
        self.rebind(self.grads, first=True)

    def rebind(self, grads: torch.Tensor, first: bool = False) -> None:
        self.grads = grads
        for (name, p), (lo, hi) in zip(self.named, self.ranges):
            n = p.numel()
            if first:
                self.params[lo:lo + n].copy_(p.data.reshape(-1).float())
                p.data = self.params[lo:lo + n].view_as(p)
            p.grad = self.grads[lo:lo + n].view_as(p)

    def set_ready_hook(self, fn: Optional[Callable[[int, int], None]]) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if fn is None:
            return
        for (_, p), (lo, hi) in zip(self.named, self.ranges):
            self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, lo=lo, hi=hi: fn(lo, hi)))

    def zero_grad(self) -> None:
        self.grads.zero_()
