import json

import torch

from b200ddl.utils import Timeline, checksum_across_ranks, timeline_from_env


def test_timeline_host_spans_and_dump(tmp_path, monkeypatch):
    p = tmp_path / "trace.json"
    monkeypatch.setenv("B200DDL_TIMELINE", str(p))
    tl = timeline_from_env(rank=0)
    assert isinstance(tl, Timeline)
    with tl.host_span("step", "host"):
        with tl.device_span("forward", "step"):  # falls back to a host span without CUDA
            sum(range(1000))
    tl.dump()
    ev = json.loads(p.read_text())["traceEvents"]
    assert {e["name"] for e in ev} == {"step", "forward"} and all(e["ph"] == "X" and e["dur"] >= 0 for e in ev)
    assert timeline_from_env(rank=3).path.endswith("trace.rank3.json")
    monkeypatch.delenv("B200DDL_TIMELINE")
    assert timeline_from_env() is None


def test_checksum_single_process():
    assert checksum_across_ranks(torch.randn(100))
