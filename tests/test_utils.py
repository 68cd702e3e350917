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


def test_usable_cpus_honours_cgroup_quota(tmp_path, monkeypatch):
    """`usable_cpus()` = min(affinity, cgroup quota): on the GPU pods cpu_count() is 128 but cpu.max grants 16."""
    import builtins

    from b200ddl.utils import cpus

    real_open = builtins.open

    def fake_open(path, *a, **k):
        if path == "/sys/fs/cgroup/cpu.max":
            p = tmp_path / "cpu.max"
            p.write_text("300000 100000\n")
            return real_open(p, *a, **k)
        return real_open(path, *a, **k)

    monkeypatch.setattr(cpus.os, "sched_getaffinity", lambda pid: set(range(64)), raising=False)
    monkeypatch.setattr(builtins, "open", fake_open)
    assert cpus.usable_cpus() == 3
    (tmp_path / "cpu.max").write_text("max 100000\n")

    def fake_open_max(path, *a, **k):
        if path == "/sys/fs/cgroup/cpu.max":
            return real_open(tmp_path / "cpu.max", *a, **k)
        return real_open(path, *a, **k)

    monkeypatch.setattr(builtins, "open", fake_open_max)
    assert cpus.usable_cpus() == 64
