"""Chrome-trace timeline (the reference's commented-out HOROVOD_TIMELINE hook, P1/03:407-409).

    B200DDL_TIMELINE=/path/trace.json python train.py        # then open in chrome://tracing / Perfetto

Device spans are measured with CUDA events recorded on the stream that runs the work (per-bucket all-reduce on the
comm stream, forward / backward / optimizer phases on the compute stream); host spans with perf_counter.  Event
timing cannot be captured inside a CUDA graph, so enabling the timeline makes the Trainer run eagerly (like
Horovod's timeline it "can incur slow down")."""
from __future__ import annotations

import json
import os
import time
from contextlib import contextmanager
from typing import List, Optional, Tuple


class Timeline:
    def __init__(self, path: str, rank: int = 0):
        self.path = path
        self.rank = rank
        self._host: List[dict] = []
        self._dev: List[Tuple[str, str, object, object]] = []
        self._t0 = time.perf_counter()
        self._anchor = None  # (cuda event, host time) to place device spans on the host clock

    @contextmanager
    def host_span(self, name: str, cat: str = "host"):
        t = time.perf_counter()
        try:
            yield
        finally:
            self._host.append({"name": name, "cat": cat, "ph": "X", "pid": self.rank, "tid": "host",
                               "ts": (t - self._t0) * 1e6, "dur": (time.perf_counter() - t) * 1e6})

    @contextmanager
    def device_span(self, name: str, cat: str = "gpu", stream=None):
        import torch

        if not torch.cuda.is_available():
            with self.host_span(name, cat):
                yield
            return
        s = stream or torch.cuda.current_stream()
        if self._anchor is None:
            a = torch.cuda.Event(enable_timing=True)
            a.record(s)
            a.synchronize()
            self._anchor = (a, time.perf_counter())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        try:
            yield
        finally:
            e1.record(s)
            self._dev.append((name, cat, e0, e1))

    def dump(self) -> str:
        import torch

        events = list(self._host)
        if self._dev:
            torch.cuda.synchronize()
            a, ta = self._anchor
            for name, cat, e0, e1 in self._dev:
                ts = (ta - self._t0) * 1e6 + a.elapsed_time(e0) * 1e3
                events.append({"name": name, "cat": cat, "ph": "X", "pid": self.rank, "tid": cat,
                               "ts": ts, "dur": e0.elapsed_time(e1) * 1e3})
        os.makedirs(os.path.dirname(os.path.abspath(self.path)) or ".", exist_ok=True)
        with open(self.path, "w") as f:
            json.dump({"traceEvents": events, "displayTimeUnit": "ms"}, f)
        return self.path


def timeline_from_env(rank: int = 0) -> Optional[Timeline]:
    p = os.environ.get("B200DDL_TIMELINE")
    if not p:
        return None
    if rank:
        root, ext = os.path.splitext(p)
        p = f"{root}.rank{rank}{ext or '.json'}"
    return Timeline(p, rank)
