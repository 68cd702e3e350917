"""Horovod-style callbacks (reference P1/03:304-322, P2/02:191-211).

The reference builds these and then forgets to pass them to `fit`; here they are honoured.
"""
from __future__ import annotations

import torch

from . import core
from ..train.callbacks import Callback


class BroadcastGlobalVariablesCallback(Callback):
    """Broadcast model (and optimizer) state from `root_rank` at the start of training so that every replica
    starts from identical weights (random init or checkpoint restore)."""

    def __init__(self, root_rank: int = 0):
        self.root_rank = root_rank
        self.done = False

    def on_train_begin(self, logs=None):
        if self.done or core.size() == 1:
            return
        self.trainer.broadcast_state(self.root_rank)
        self.done = True


class MetricAverageCallback(Callback):
    """Average epoch-end metrics over all ranks (must run before metric-driven callbacks such as
    ReduceLROnPlateau - the ordering note at P1/03:310-313)."""

    def on_epoch_end(self, epoch, logs=None):
        if logs is None or core.size() == 1:
            return
        keys = sorted(k for k, v in logs.items() if isinstance(v, (int, float)))
        if not keys:
            return
        t = torch.tensor([float(logs[k]) for k in keys], dtype=torch.float64, device=core.device())
        t = core.allreduce(t.float() if t.device.type == "cuda" else t, average=True)
        for k, v in zip(keys, t.tolist()):
            logs[k] = v


class LearningRateWarmupCallback(Callback):
    """Ramp the LR from `initial_lr / size` to `initial_lr` over `warmup_epochs` (Goyal et al. 1706.02677, cited at
    P1/03:315-318), per batch."""

    def __init__(self, initial_lr: float, warmup_epochs: int = 5, steps_per_epoch: int = None, verbose: int = 0):
        self.initial_lr = float(initial_lr)
        self.warmup_epochs = warmup_epochs
        self.steps_per_epoch = steps_per_epoch
        self.verbose = verbose
        self._epoch = 0

    def on_epoch_begin(self, epoch, logs=None):
        self._epoch = epoch

    def on_train_batch_begin(self, batch, logs=None):
        if self._epoch >= self.warmup_epochs:
            return
        spe = self.steps_per_epoch or self.trainer.steps_per_epoch or 1
        progress = (self._epoch * spe + batch) / float(max(1, self.warmup_epochs * spe))
        n = core.size()
        self.trainer.optimizer.learning_rate = self.initial_lr * (1.0 / n + progress * (1.0 - 1.0 / n))

    def on_epoch_end(self, epoch, logs=None):
        if epoch == self.warmup_epochs - 1:
            self.trainer.optimizer.learning_rate = self.initial_lr
            if self.verbose and core.rank() == 0:
                print(f"Epoch {epoch + 1}: finished gradual learning rate warmup to {self.initial_lr:g}.")
