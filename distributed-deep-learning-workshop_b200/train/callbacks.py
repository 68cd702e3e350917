"""Keras-style callbacks (reference: ReduceLROnPlateau P1/03:321, ModelCheckpoint P2/02:206-211, EarlyStopping
P2/03:397-401 - all constructed in the reference and then never passed to `fit`).  Here
`Trainer.fit(callbacks=[...])` honours them."""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional


class Callback:
    trainer = None

    def set_trainer(self, trainer) -> None:
        self.trainer = trainer

    def on_train_begin(self, logs=None): ...
    def on_train_end(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...
    def on_train_batch_begin(self, batch, logs=None): ...
    def on_train_batch_end(self, batch, logs=None): ...


class CallbackList:
    def __init__(self, callbacks: Optional[List[Callback]], trainer):
        self.callbacks = list(callbacks or [])
        for c in self.callbacks:
            c.set_trainer(trainer)

    def call(self, hook: str, *args, **kw) -> None:
        for c in self.callbacks:
            getattr(c, hook)(*args, **kw)


class History(Callback):
    """`history['val_loss'][-1]` is the reference's in-process metric store (P1/03:369,375)."""

    def __init__(self):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


class LambdaCallback(Callback):
    def __init__(self, on_epoch_end: Optional[Callable] = None, on_train_batch_end: Optional[Callable] = None,
                 on_train_begin: Optional[Callable] = None, on_train_end: Optional[Callable] = None):
        self._ee, self._be, self._tb, self._te = on_epoch_end, on_train_batch_end, on_train_begin, on_train_end

    def on_epoch_end(self, epoch, logs=None):
        if self._ee:
            self._ee(epoch, logs)

    def on_train_batch_end(self, batch, logs=None):
        if self._be:
            self._be(batch, logs)

    def on_train_begin(self, logs=None):
        if self._tb:
            self._tb(logs)

    def on_train_end(self, logs=None):
        if self._te:
            self._te(logs)


def _better(mode: str, monitor: str):
    if mode == "auto":
        mode = "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"
    return (lambda a, b, d: a > b + d) if mode == "max" else (lambda a, b, d: a < b - d)


class EarlyStopping(Callback):
    def __init__(self, monitor: str = "val_loss", min_delta: float = 0.0, patience: int = 0, mode: str = "auto",
                 restore_best_weights: bool = False, verbose: int = 0):
        self.monitor, self.min_delta, self.patience = monitor, abs(min_delta), patience
        self.better = _better(mode, monitor)
        self.restore_best_weights = restore_best_weights
        self.verbose = verbose
        self.best = None
        self.wait = 0
        self.stopped_epoch = None
        self._best_state = None

    def on_train_begin(self, logs=None):
        self.best, self.wait, self.stopped_epoch = None, 0, None

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None or (isinstance(cur, float) and math.isnan(cur)):
            return
        if self.best is None or self.better(cur, self.best, self.min_delta):
            self.best, self.wait = cur, 0
            if self.restore_best_weights:
                self._best_state = self.trainer.get_weights()
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.trainer.stop_training = True
                if self.restore_best_weights and self._best_state is not None:
                    self.trainer.set_weights(self._best_state)
                if self.verbose:
                    print(f"Epoch {epoch + 1}: early stopping")


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor: str = "val_loss", factor: float = 0.1, patience: int = 10, min_delta: float = 1e-4,
                 cooldown: int = 0, min_lr: float = 0.0, mode: str = "auto", verbose: int = 0):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0")
        self.monitor, self.factor, self.patience = monitor, factor, patience
        self.min_delta, self.cooldown, self.min_lr, self.verbose = min_delta, cooldown, min_lr, verbose
        self.better = _better(mode, monitor)
        self.best = None
        self.wait = 0
        self.cooldown_counter = 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        opt = self.trainer.optimizer
        logs["lr"] = float(opt.learning_rate)
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if self.best is None or self.better(cur, self.best, self.min_delta):
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = float(opt.learning_rate)
                if old > self.min_lr:
                    new = max(old * self.factor, self.min_lr)
                    opt.learning_rate = new
                    if self.verbose:
                        print(f"Epoch {epoch + 1}: ReduceLROnPlateau reducing learning rate to {new:g}.")
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class ModelCheckpoint(Callback):
    """Per-epoch checkpoint; ``filepath`` may contain ``{epoch}`` and metric names.  In a distributed job only
    rank 0 writes ("Save checkpoints only on worker 0 to prevent conflicts between workers", P2/02:206-208)."""

    def __init__(self, filepath: str, save_weights_only: bool = True, monitor: str = "val_loss",
                 save_best_only: bool = False, mode: str = "auto", verbose: int = 0):
        self.filepath, self.save_weights_only = filepath, save_weights_only
        self.monitor, self.save_best_only, self.verbose = monitor, save_best_only, verbose
        self.better = _better(mode, monitor)
        self.best = None
        self.saved: List[str] = []

    def on_epoch_end(self, epoch, logs=None):
        from ..parallel import core

        if core.rank() != 0:
            return
        logs = logs or {}
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None or (self.best is not None and not self.better(cur, self.best, 0.0)):
                return
            self.best = cur
        path = self.filepath.format(epoch=epoch + 1, **{k: v for k, v in logs.items() if isinstance(v, (int, float))})
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        self.trainer.save_weights(path) if self.save_weights_only else self.trainer.save(path)
        self.saved.append(path)
        if self.verbose:
            print(f"Epoch {epoch + 1}: saving model to {path}")
