"""Trainer: the Keras `compile / fit / evaluate / predict` surface of the reference (P1/02:194-215,
P1/03:204-234,325-358, P2/03:206) on top of two execution backends:

* **engine backend** - `models.ResNet50Engine`: hand-scheduled sm_100a kernels, whole step in one CUDA graph,
  gradients born inside the (symmetric) flat buffer that `DistributedOptimizer` all-reduces while backward runs;
* **module backend** - any `torch.nn.Module` (e.g. the reference's frozen MobileNetV2 + Dense head, or CPU runs):
  autograd computes gradients into the same kind of flat buffer (`FlatParams`), so optimizers, callbacks and the
  distributed wrapper are shared.

Datasets are iterables of ``(images uint8 [B,H,W,3], labels int64 [B])`` (see `loader.Converter.make_dataset`).
"""
from __future__ import annotations

import time
from collections import deque
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .callbacks import Callback, CallbackList, History
from .flat import FlatParams
from ..parallel import core as dist_core


def _is_engine(model) -> bool:
    return hasattr(model, "param_specs") and hasattr(model, "forward") and hasattr(model, "bind_grad_buffer")


class _EngineBackend:
    def __init__(self, engine, optimizer, use_graph: bool):
        from ..models.resnet_engine import EngineTrainStep, EngineEvalStep

        self.engine = engine
        self.optimizer = optimizer
        self.use_graph = use_graph
        # Both step objects are created on first use: a Trainer that only ever predicts (pyfunc serving, load_model) builds
        # the engine for INFERENCE - BatchNorm folded into the convolution epilogues, no gradient / backward buffers - and
        # only a later train_batch() rebuilds it for training.
        self._step = None
        self._eval_step = None
        self.inference_only = False
        self.batch = engine.batch
        self._pending = deque()
        # one pinned result slot per IN-FLIGHT step: a slot goes back to the free list only after its value has been
        # read, so queueing many batches before popping (evaluate) can never overwrite an unread result
        self._free = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(4)]

    @property
    def step(self):
        if self._step is None:
            from ..models.resnet_engine import EngineTrainStep

            self._step = EngineTrainStep(self.engine, self.optimizer, use_graph=self.use_graph)
            self._eval_step = None  # the engine was (re)built for training: eval plans / graph must be re-captured
        return self._step

    @property
    def eval_step(self):
        if self._eval_step is None:
            from ..models.resnet_engine import EngineEvalStep

            if self._step is None and not self.inference_only:
                _ = self.step  # training-capable trainer: evaluate on the training build (same buffers)
            self._eval_step = EngineEvalStep(self.engine, use_graph=self.use_graph)
        return self._eval_step

    def _enqueue_result(self) -> None:
        buf = self._free.pop() if self._free else torch.zeros(2, dtype=torch.float32).pin_memory()
        buf.copy_(self.engine.stats, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, buf))

    def pop_results(self, keep: int = 0) -> List[Tuple[float, float]]:
        out = []
        while len(self._pending) > keep:
            ev, buf = self._pending.popleft()
            ev.synchronize()
            out.append((float(buf[0]) / self.batch, float(buf[1]) / self.batch))
            self._free.append(buf)
        return out

    def train_batch(self, x, y) -> None:
        self.step.load(x, y)
        self.step.run()
        self._enqueue_result()

    def eval_batch(self, x, y) -> None:
        step = self.eval_step  # (builds the engine on first use)
        self.engine.set_input(x, y)
        step.run()
        self._enqueue_result()

    def predict_batch(self, x) -> torch.Tensor:
        step = self.eval_step
        self.engine.set_input(x, None)
        step.run()
        return self.engine.logits.clone()

    def predict_stream(self, arr: torch.Tensor) -> torch.Tensor:
        """Logits [n, classes] (host, fp32) for uint8 images [n, H, W, 3] - the whole array in ONE pipeline:
        batch k+1 is copied host->device on a side stream into one of two staging buffers while the eval graph of batch
        k runs; the logits of every batch go to one pinned result array with async copies; a single sync at the end.
        A pinned `arr` (what the pyfunc scoring workers provide) makes the H2D copies truly asynchronous."""
        e, B = self.engine, self.batch
        step = self.eval_step  # builds the engine (for inference when this trainer never trained) before x_u8 is touched
        n = int(arr.shape[0])
        K = e.num_classes
        out = torch.empty((n, K), dtype=torch.float32).pin_memory() if n else torch.empty((0, K))
        if n == 0:
            return out
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(e.x_u8) for _ in range(2)]
            self._stage_free = [None, None]   # event: the D2D out of the staging buffer has run
            self._copy_stream = torch.cuda.Stream(device=e.device)
        cur = torch.cuda.current_stream()
        self._copy_stream.wait_stream(cur)
        nb = (n + B - 1) // B
        flat = arr.reshape(n, -1)

        def stage(k):
            s = k & 1
            m = min(B, n - k * B)
            with torch.cuda.stream(self._copy_stream):
                if self._stage_free[s] is not None:
                    self._copy_stream.wait_event(self._stage_free[s])
                self._stage[s].view(B, -1)[:m].copy_(flat[k * B:k * B + m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            return ev

        ready = stage(0)
        for k in range(nb):
            s = k & 1
            m = min(B, n - k * B)
            nxt = stage(k + 1) if k + 1 < nb else None
            cur.wait_event(ready)
            e.x_u8.copy_(self._stage[s], non_blocking=True)   # D2D, ~15 us
            done = torch.cuda.Event()
            done.record()
            self._stage_free[s] = done
            step.run()
            out[k * B:k * B + m].copy_(e.logits[:m], non_blocking=True)
            ready = nxt
        cur.synchronize()
        return out

    def state_tensors(self) -> List[torch.Tensor]:
        e = self.engine
        if not self.inference_only:
            _ = self.step  # optimizer state exists (and weights may have moved to symmetric memory) only once attached
        opt = self.optimizer.opt if hasattr(self.optimizer, "opt") else self.optimizer
        return [e.params, e.running] + list(opt.state.values())

    def after_state_change(self) -> None:
        self.engine.sync_weights()

    def state_dict(self):
        return self.engine.state_dict()

    def load_state_dict(self, sd) -> None:
        self.engine.load_state_dict(sd)


class _ModuleBackend:
    def __init__(self, module: torch.nn.Module, optimizer, preprocess, device: torch.device):
        self.module = module.to(device)
        self.device = device
        self.optimizer = optimizer
        self.preprocess = preprocess
        self.flat = FlatParams(self.module)
        self._dist = hasattr(optimizer, "on_grads_ready")
        if self._dist:
            grads = optimizer.allocate_grads(self.flat.numel, device)
            self.flat.rebind(grads)
            optimizer.attach(self.flat.params, self.flat.ranges, None, grads)
            self.flat.set_ready_hook(optimizer.on_grads_ready)
        else:
            optimizer.attach(self.flat.params, self.flat.grads, None)
        self._results: deque = deque()
        self.batch = None

    def _prep(self, x) -> torch.Tensor:
        x = torch.as_tensor(x).to(self.device, non_blocking=True)
        return self.preprocess(x)

    def train_batch(self, x, y) -> None:
        self.module.train()
        y = torch.as_tensor(y).to(self.device, non_blocking=True)
        self.optimizer.begin_step()
        self.flat.zero_grad()
        if self._dist:
            self.optimizer.start_backward()
        logits = self.module(self._prep(x)).float()
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        self.optimizer.step()
        acc = (logits.argmax(1) == y).float().mean()
        self._results.append(torch.stack([loss.detach(), acc]))

    @torch.no_grad()
    def eval_batch(self, x, y) -> None:
        self.module.eval()
        y = torch.as_tensor(y).to(self.device, non_blocking=True)
        logits = self.module(self._prep(x)).float()
        loss = torch.nn.functional.cross_entropy(logits, y)
        acc = (logits.argmax(1) == y).float().mean()
        self._results.append(torch.stack([loss, acc]))

    @torch.no_grad()
    def predict_batch(self, x) -> torch.Tensor:
        self.module.eval()
        return self.module(self._prep(x)).float()

    def pop_results(self, keep: int = 0) -> List[Tuple[float, float]]:
        out = []
        while len(self._results) > keep:
            t = self._results.popleft().tolist()
            out.append((t[0], t[1]))
        return out

    def state_tensors(self) -> List[torch.Tensor]:
        opt = self.optimizer.opt if hasattr(self.optimizer, "opt") else self.optimizer
        bufs = [b for b in self.module.buffers() if b.dtype.is_floating_point]
        return [self.flat.params] + bufs + list(opt.state.values())

    def after_state_change(self) -> None:
        pass

    def state_dict(self):
        return {k: v.detach().cpu().clone() for k, v in self.module.state_dict().items()}

    def load_state_dict(self, sd) -> None:
        own = self.module.state_dict()
        for k, v in sd.items():
            own[k].copy_(v)


class Trainer:
    def __init__(self, model, preprocess=None, device: Optional[torch.device] = None, use_graph: bool = True):
        self.model = model
        self.device = torch.device(device) if device is not None else (
            getattr(model, "device", None) or dist_core.device())
        if preprocess is None:
            from ..models.preprocess import preprocess_tensor

            preprocess = preprocess_tensor
        self._preprocess = preprocess
        self.use_graph = use_graph
        self.optimizer = None
        self.backend = None
        self.metrics_names = ["loss"]
        self.stop_training = False
        self.steps_per_epoch: Optional[int] = None
        self.history: Optional[History] = None
        self.extra_callbacks: List[Callback] = []  # e.g. tracking.autolog()

    # ------------------------------------------------------------------------------------------------ compile
    def compile(self, optimizer, loss: str = "sparse_categorical_crossentropy", metrics: Sequence[str] = ("accuracy",),
                **_ignored) -> "Trainer":
        """`model.compile(optimizer=..., loss=SparseCategoricalCrossentropy(from_logits=True), metrics=['accuracy'])`"""
        loss_name = loss if isinstance(loss, str) else getattr(loss, "name", "sparse_categorical_crossentropy")
        if "sparse_categorical_crossentropy" not in loss_name:
            raise ValueError("only sparse categorical cross-entropy from logits is implemented (what the reference uses)")
        self.optimizer = optimizer
        self.metrics_names = ["loss"] + [m if isinstance(m, str) else getattr(m, "name", "metric") for m in metrics]
        if _is_engine(self.model):
            self.backend = _EngineBackend(self.model, optimizer, self.use_graph)
        else:
            self.backend = _ModuleBackend(self.model, optimizer, self._preprocess, self.device)
        return self

    # ------------------------------------------------------------------------------------------------ fit
    def fit(self, x: Iterable, steps_per_epoch: Optional[int] = None, epochs: int = 1, verbose: int = 1,
            validation_data: Optional[Iterable] = None, validation_steps: Optional[int] = None,
            callbacks: Optional[List[Callback]] = None, initial_epoch: int = 0, **_ignored) -> History:
        if self.backend is None:
            raise RuntimeError("call compile() before fit()")
        if steps_per_epoch is None:
            steps_per_epoch = len(x)  # finite datasets only
        self.steps_per_epoch = steps_per_epoch
        self.history = History()
        cbs = CallbackList([*(callbacks or []), *self.extra_callbacks, self.history], self)
        self.stop_training = False
        is_chief = dist_core.rank() == 0
        it = iter(x)
        cbs.call("on_train_begin", {})
        for epoch in range(initial_epoch, epochs):
            cbs.call("on_epoch_begin", epoch, {})
            t0 = time.perf_counter()
            tot_loss = tot_acc = 0.0
            n_seen = 0
            for b in range(steps_per_epoch):
                cbs.call("on_train_batch_begin", b, {})
                try:
                    xb, yb = next(it)
                except StopIteration:
                    it = iter(x)
                    try:
                        xb, yb = next(it)
                    except StopIteration:
                        raise ValueError("fit(): the dataset yields no batches (fewer rows than batch_size?)") from None
                self.backend.train_batch(xb, yb)
                for l, a in self.backend.pop_results(keep=1):  # lag one step: keeps H2D/compute overlapped
                    tot_loss += l
                    tot_acc += a
                    n_seen += 1
                cbs.call("on_train_batch_end", b, {})
            for l, a in self.backend.pop_results(keep=0):
                tot_loss += l
                tot_acc += a
                n_seen += 1
            dt = time.perf_counter() - t0
            logs: Dict[str, float] = {"loss": tot_loss / max(n_seen, 1)}
            if "accuracy" in self.metrics_names:
                logs["accuracy"] = tot_acc / max(n_seen, 1)
            bsz = self._batch_size(xb)
            logs["images_per_sec"] = steps_per_epoch * bsz * dist_core.size() / max(dt, 1e-9)
            if validation_data is not None:
                vl = self.evaluate(validation_data, steps=validation_steps, verbose=0)
                logs["val_loss"] = vl[0]
                if len(vl) > 1:
                    logs["val_accuracy"] = vl[1]
            cbs.call("on_epoch_end", epoch, logs)
            if verbose and is_chief:
                msg = " - ".join(f"{k}: {v:.4f}" for k, v in logs.items() if k != "images_per_sec")
                print(f"Epoch {epoch + 1}/{epochs} - {dt:.1f}s - {msg} - {logs['images_per_sec']:.0f} img/s", flush=True)
            if self.stop_training:
                break
        cbs.call("on_train_end", {})
        step = getattr(self.backend, "step", None)
        if step is not None and getattr(step, "timeline", None) is not None:
            path = step.dump_timeline()
            if is_chief:
                print(f"timeline written to {path}")
        return self.history

    @staticmethod
    def _batch_size(xb) -> int:
        return int(xb.shape[0]) if hasattr(xb, "shape") and len(xb.shape) == 4 else int(len(xb))

    # ------------------------------------------------------------------------------------------------ evaluate
    def evaluate(self, x: Iterable, steps: Optional[int] = None, verbose: int = 0, **_ignored) -> List[float]:
        if self.backend is None:
            raise RuntimeError("call compile() before evaluate()")
        tot_loss = tot_acc = 0.0
        n = 0
        it = iter(x)
        k = 0
        while steps is None or k < steps:
            try:
                xb, yb = next(it)
            except StopIteration:
                if steps is None:
                    break
                it = iter(x)
                try:
                    xb, yb = next(it)
                except StopIteration:
                    break  # empty dataset
            self.backend.eval_batch(xb, yb)
            k += 1
            for l, a in self.backend.pop_results(keep=2):  # bounded queue; lag keeps H2D / compute overlapped
                tot_loss += l
                tot_acc += a
                n += 1
        for l, a in self.backend.pop_results(keep=0):
            tot_loss += l
            tot_acc += a
            n += 1
        out = [tot_loss / max(n, 1)]
        if "accuracy" in self.metrics_names:
            out.append(tot_acc / max(n, 1))
        if verbose and dist_core.rank() == 0:
            print(" - ".join(f"{m}: {v:.4f}" for m, v in zip(self.metrics_names, out)))
        return out

    # ------------------------------------------------------------------------------------------------ predict
    def predict(self, x, batch_size: int = 32, **_ignored) -> np.ndarray:
        """`model.predict(np.array, batch_size)` -> logits [n, classes] (reference P2/03:206)."""
        if self.backend is None:
            self._compile_for_inference()
        arr = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        n = arr.shape[0]
        if (hasattr(self.backend, "predict_stream") and arr.dtype == torch.uint8 and arr.dim() == 4
                and tuple(arr.shape[1:]) == (self.backend.engine.image_size, self.backend.engine.image_size, 3)):
            return self.backend.predict_stream(arr).numpy()
        fixed = getattr(self.backend, "batch", None)
        bs = fixed or batch_size
        outs = []
        for i in range(0, n, bs):
            chunk = arr[i:i + bs]
            m = chunk.shape[0]
            if fixed and m < bs:  # static-shape engine: pad the tail batch
                pad = torch.zeros((bs - m, *chunk.shape[1:]), dtype=chunk.dtype)
                chunk = torch.cat([chunk, pad.to(chunk.device)], 0)
            outs.append(self.backend.predict_batch(chunk)[:m].cpu())
        return torch.cat(outs, 0).numpy() if outs else np.zeros((0, 0), np.float32)

    def _compile_for_inference(self) -> None:
        from .. import optim

        self.compile(optim.SGD(0.0))
        if hasattr(self.backend, "inference_only"):
            self.backend.inference_only = True  # engine backend: build with the fused inference epilogues

    # ------------------------------------------------------------------------------------------------ state
    def broadcast_state(self, root: int = 0) -> None:
        """K2: identical weights / optimizer state / BN statistics on every rank."""
        for t in self.backend.state_tensors():
            dist_core.broadcast(t, root)
        self.backend.after_state_change()

    def get_weights(self):
        return self.backend.state_dict()

    def set_weights(self, sd) -> None:
        self.backend.load_state_dict(sd)

    def save_weights(self, path: str) -> None:
        torch.save(self.backend.state_dict(), path)

    def load_weights(self, path: str) -> None:
        self.backend.load_state_dict(torch.load(path, map_location="cpu"))

    def save(self, path: str) -> None:
        """Weights + optimizer state + step counter.  With `DistributedOptimizer(fused_update=True)` the momentum is sharded
        over ranks: call `optimizer.consolidate_state()` on EVERY rank first (it is a collective; `save` usually runs on
        rank 0 only and therefore cannot do it itself) - an unconsolidated state is refused, not silently truncated."""
        if getattr(self.optimizer, "state_is_sharded", False):
            raise RuntimeError("the optimizer state is sharded over ranks (fused_update): call "
                               "optimizer.consolidate_state() on every rank before Trainer.save(), or use save_weights()")
        opt = self.optimizer.opt if hasattr(self.optimizer, "opt") else self.optimizer
        torch.save({"weights": self.backend.state_dict(), "optimizer": opt.state_dict() if opt is not None else None,
                    "arch": getattr(self.model, "arch", type(self.model).__name__)}, path)

    def load(self, path: str, load_optimizer: bool = True) -> "Trainer":
        """Restore what `save()` (or `save_weights()` / `ModelCheckpoint`) wrote: weights and, when present and wanted,
        the optimizer state and step counter.  Resume recipe (the reference only alludes to it at
        P1/03:305-307): `compile(...)`, `load(path)` on every rank or on rank 0 followed by
        `BroadcastGlobalVariablesCallback(0)` / `broadcast_state(0)`, then `fit(..., initial_epoch=k)`."""
        if self.backend is None:
            raise RuntimeError("call compile() before load()")
        ck = torch.load(path, map_location="cpu")
        weights = ck["weights"] if isinstance(ck, dict) and "weights" in ck else ck
        self.backend.load_state_dict(weights)
        if load_optimizer and isinstance(ck, dict) and ck.get("optimizer") is not None and self.optimizer is not None:
            opt = self.optimizer.opt if hasattr(self.optimizer, "opt") else self.optimizer
            if ck["optimizer"].get("name") != getattr(opt, "name", None):
                raise ValueError(f"checkpoint holds {ck['optimizer'].get('name')} state, the trainer was compiled with "
                                 f"{getattr(opt, 'name', type(opt).__name__)}")
            opt.load_state_dict(ck["optimizer"])
        return self

    def summary(self) -> str:
        if _is_engine(self.model):
            n = self.model.num_parameters()
            tr = self.model.trainable_parameters() if hasattr(self.model, "trainable_parameters") else n
            lines = [f"Model: {self.model.arch} (B200 engine)", f"Total params: {n:,}", f"Trainable params: {tr:,}"]
            if tr != n:
                lines.append(f"Non-trainable params: {n - tr:,}")
        else:
            tot = sum(p.numel() for p in self.model.parameters())
            tr = sum(p.numel() for p in self.model.parameters() if p.requires_grad)
            lines = [f"Model: {type(self.model).__name__}", f"Total params: {tot:,}", f"Trainable params: {tr:,}",
                     f"Non-trainable params: {tot - tr:,}"]
        s = "\n".join(lines)
        print(s)
        return s
