"""Experiment tracking, model registry and model logging - the MLflow surface the reference uses
on a local file-backed store.  Verb names and semantics follow MLflow so notebook code ports 1:1::

    set_tracking_uri, set_experiment, start_run(run_name|run_id|nested|experiment_id), active_run, end_run,
    log_param(s), log_metric(s), log_dict, log_artifact, set_tag, search_runs(filter_string, order_by),
    register_model, MlflowClient().transition_model_version_stage, autolog,
    tracking.keras.log_model / load_model  (the reference's mlflow.keras.*, P1/03:373,438)

Store layout::  <uri>/experiments.json, <uri>/<exp_id>/<run_id>/{meta.json, params.json, metrics.jsonl, tags.json,
artifacts/...}, <uri>/models/<name>/{meta.json}.   URIs: ``runs:/<run_id>/<path>``, ``models:/<name>/<stage|version>``.
Worker processes started by `parallel.Runner` inherit the URI through ``B200DDL_TRACKING_URI`` (the reference ships
DATABRICKS_HOST/TOKEN to workers for the same reason, P1/03:286-288).
"""
from __future__ import annotations

import json
import os
import shutil
import threading
import time
import uuid
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import pandas as pd

_state = threading.local()
_lock = threading.RLock()
_uri: Optional[str] = None
_experiment_id: Optional[str] = None
_autolog_enabled = False


# ------------------------------------------------------------------------------------------------ store helpers
def set_tracking_uri(uri: str) -> None:
    global _uri
    if uri in ("databricks", None):  # reference calls set_tracking_uri("databricks"); keep the configured store
        uri = _uri or os.environ.get("B200DDL_TRACKING_URI") or os.path.join(os.getcwd(), "mlruns")
    if uri.startswith("file:"):
        uri = uri[5:]
    _uri = os.path.abspath(uri)
    os.makedirs(_uri, exist_ok=True)
    os.environ["B200DDL_TRACKING_URI"] = _uri


def get_tracking_uri() -> str:
    if _uri is None:
        set_tracking_uri(os.environ.get("B200DDL_TRACKING_URI") or os.path.join(os.getcwd(), "mlruns"))
    return _uri


def _read_json(path: str, default):
    try:
        with open(path) as f:
            return json.load(f)
    except (FileNotFoundError, json.JSONDecodeError):
        return default


def _write_json(path: str, obj) -> None:
    tmp = f"{path}.{uuid.uuid4().hex[:6]}.tmp"
    with open(tmp, "w") as f:
        json.dump(obj, f)
    os.replace(tmp, path)


def _experiments() -> Dict[str, dict]:
    return _read_json(os.path.join(get_tracking_uri(), "experiments.json"), {})


def set_experiment(name: str) -> SimpleNamespace:
    """`mlflow.set_experiment('/Users/<user>/distributed_dl_workshop')` (reference P2/01:221)."""
    global _experiment_id
    with _lock:
        exps = _experiments()
        for eid, e in exps.items():
            if e["name"] == name:
                _experiment_id = eid
                return SimpleNamespace(experiment_id=eid, name=name)
        eid = str(len(exps))
        exps[eid] = {"name": name, "created": time.time()}
        _write_json(os.path.join(get_tracking_uri(), "experiments.json"), exps)
        os.makedirs(os.path.join(get_tracking_uri(), eid), exist_ok=True)
        _experiment_id = eid
        return SimpleNamespace(experiment_id=eid, name=name)


def _current_experiment() -> str:
    global _experiment_id
    if _experiment_id is None:
        env = os.environ.get("B200DDL_EXPERIMENT_ID")
        if env is not None:
            _experiment_id = env
        else:
            set_experiment("Default")
    return _experiment_id


def _find_run_dir(run_id: str) -> str:
    root = get_tracking_uri()
    for eid in os.listdir(root):
        d = os.path.join(root, eid, run_id)
        if os.path.isdir(d):
            return d
    raise KeyError(f"run {run_id} not found in {root}")


# ------------------------------------------------------------------------------------------------ runs
class Run:
    def __init__(self, run_dir: str):
        self._dir = run_dir
        meta = _read_json(os.path.join(run_dir, "meta.json"), {})
        self.info = SimpleNamespace(run_id=meta.get("run_id"), run_uuid=meta.get("run_id"),
                                    experiment_id=meta.get("experiment_id"), run_name=meta.get("run_name"),
                                    status=meta.get("status"), start_time=meta.get("start_time"),
                                    end_time=meta.get("end_time"), artifact_uri=os.path.join(run_dir, "artifacts"))

    @property
    def data(self) -> SimpleNamespace:
        return SimpleNamespace(params=_read_json(os.path.join(self._dir, "params.json"), {}),
                               metrics=_latest_metrics(self._dir), tags=_read_json(os.path.join(self._dir, "tags.json"), {}))

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        end_run("FAILED" if et is not None else "FINISHED")
        return False


def _stack() -> List[Run]:
    if not hasattr(_state, "stack"):
        _state.stack = []
    return _state.stack


def start_run(run_id: Optional[str] = None, experiment_id: Optional[str] = None, run_name: Optional[str] = None,
              nested: bool = False, tags: Optional[dict] = None) -> Run:
    st = _stack()
    if st and not nested and run_id is None:
        raise RuntimeError(f"run {st[-1].info.run_id} is already active; use nested=True or end_run() first")
    if run_id is not None:  # resume (rank-0 worker re-opens the driver's run: reference P1/03:363)
        run = Run(_find_run_dir(run_id))
        st.append(run)
        return run
    eid = str(experiment_id) if experiment_id is not None else _current_experiment()
    rid = uuid.uuid4().hex
    d = os.path.join(get_tracking_uri(), eid, rid)
    os.makedirs(os.path.join(d, "artifacts"), exist_ok=True)
    all_tags = dict(tags or {})
    if run_name:
        all_tags["mlflow.runName"] = run_name
    if nested and st:
        all_tags["mlflow.parentRunId"] = st[-1].info.run_id
    _write_json(os.path.join(d, "meta.json"), {"run_id": rid, "experiment_id": eid, "run_name": run_name,
                                               "status": "RUNNING", "start_time": time.time(), "end_time": None})
    _write_json(os.path.join(d, "tags.json"), all_tags)
    _write_json(os.path.join(d, "params.json"), {})
    run = Run(d)
    st.append(run)
    return run


def active_run() -> Optional[Run]:
    st = _stack()
    return st[-1] if st else None


def end_run(status: str = "FINISHED") -> None:
    st = _stack()
    if not st:
        return  # the reference calls mlflow.end_run() after a `with` block (no-op)
    run = st.pop()
    meta_p = os.path.join(run._dir, "meta.json")
    meta = _read_json(meta_p, {})
    meta.update(status=status, end_time=time.time())
    _write_json(meta_p, meta)


def get_run(run_id: str) -> Run:
    return Run(_find_run_dir(run_id))


def _active_dir() -> str:
    run = active_run()
    if run is None:
        run = start_run()
    return run._dir


# ------------------------------------------------------------------------------------------------ logging
def log_param(key: str, value: Any) -> None:
    with _lock:
        p = os.path.join(_active_dir(), "params.json")
        d = _read_json(p, {})
        d[key] = str(value)
        _write_json(p, d)


def log_params(params: Dict[str, Any]) -> None:
    for k, v in params.items():
        log_param(k, v)


def log_metric(key: str, value: float, step: Optional[int] = None) -> None:
    with _lock:
        with open(os.path.join(_active_dir(), "metrics.jsonl"), "a") as f:
            f.write(json.dumps({"key": key, "value": float(value), "step": step, "ts": time.time()}) + "\n")


def log_metrics(metrics: Dict[str, float], step: Optional[int] = None) -> None:
    for k, v in metrics.items():
        log_metric(k, v, step)


def set_tag(key: str, value: Any) -> None:
    with _lock:
        p = os.path.join(_active_dir(), "tags.json")
        d = _read_json(p, {})
        d[key] = str(value)
        _write_json(p, d)


def set_tags(tags: Dict[str, Any]) -> None:
    for k, v in tags.items():
        set_tag(k, v)


def log_text(text: str, artifact_file: str) -> None:
    """`mlflow.log_text('...', 'notes.txt')`."""
    p = os.path.join(_active_dir(), "artifacts", artifact_file)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "w") as f:
        f.write(text)


def log_artifacts(local_dir: str, artifact_path: Optional[str] = None) -> None:
    """`mlflow.log_artifacts(dir)`: the CONTENTS of `local_dir` go under `artifact_path` (not the directory itself)."""
    dst = os.path.join(_active_dir(), "artifacts", artifact_path or "")
    os.makedirs(dst, exist_ok=True)
    shutil.copytree(local_dir, dst, dirs_exist_ok=True)


def get_experiment_by_name(name: str) -> Optional[SimpleNamespace]:
    for eid, e in _experiments().items():
        if e["name"] == name:
            return SimpleNamespace(experiment_id=eid, name=name)
    return None


def log_dict(dictionary: dict, artifact_file: str) -> None:
    """`mlflow.log_dict({'img_height':..,'img_width':..}, 'img_params_dict.json')` (reference P2/03:284-285)."""
    p = os.path.join(_active_dir(), "artifacts", artifact_file)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "w") as f:
        json.dump(dictionary, f)


def log_artifact(local_path: str, artifact_path: Optional[str] = None) -> None:
    dst = os.path.join(_active_dir(), "artifacts", artifact_path or "")
    os.makedirs(dst, exist_ok=True)
    if os.path.isdir(local_path):
        shutil.copytree(local_path, os.path.join(dst, os.path.basename(local_path)), dirs_exist_ok=True)
    else:
        shutil.copy2(local_path, dst)


def get_artifact_uri(artifact_path: str = "") -> str:
    return os.path.join(_active_dir(), "artifacts", artifact_path)


def _latest_metrics(run_dir: str) -> Dict[str, float]:
    out: Dict[str, float] = {}
    try:
        with open(os.path.join(run_dir, "metrics.jsonl")) as f:
            for line in f:
                r = json.loads(line)
                out[r["key"]] = r["value"]
    except FileNotFoundError:
        pass
    return out


def metric_history(run_id: str, key: str) -> List[float]:
    vals = []
    try:
        with open(os.path.join(_find_run_dir(run_id), "metrics.jsonl")) as f:
            for line in f:
                r = json.loads(line)
                if r["key"] == key:
                    vals.append(r["value"])
    except FileNotFoundError:
        pass
    return vals


# ------------------------------------------------------------------------------------------------ search
def search_runs(experiment_ids: Optional[List[str]] = None, filter_string: str = "",
                order_by: Optional[List[str]] = None, max_results: int = 100000) -> pd.DataFrame:
    """`mlflow.search_runs(filter_string='tags.mlflow.parentRunId = "<id>"', order_by=['metrics.accuracy DESC'])`
    (reference P2/01:257-258).  Supports `=` / `!=` on tags./params. and comparisons on metrics., AND-combined;
    a missing order-by metric sorts last instead of raising."""
    import re

    root = get_tracking_uri()
    eids = [str(e) for e in experiment_ids] if experiment_ids else [_current_experiment()]
    conds = []
    for clause in [c for c in re.split(r"(?i)\s+and\s+", filter_string.strip()) if c]:
        m = re.match(r"""^\s*(tags|params|metrics|attributes)\.[`"]?([\w.\-/ ]+?)[`"]?\s*(=|!=|>=|<=|>|<)\s*['"]?([^'"]*)['"]?\s*$""", clause)
        if not m:
            raise ValueError(f"cannot parse filter clause {clause!r}")
        conds.append(m.groups())
    rows = []
    for eid in eids:
        edir = os.path.join(root, eid)
        if not os.path.isdir(edir):
            continue
        for rid in os.listdir(edir):
            d = os.path.join(edir, rid)
            meta = _read_json(os.path.join(d, "meta.json"), None)
            if not meta:
                continue
            params = _read_json(os.path.join(d, "params.json"), {})
            tags = _read_json(os.path.join(d, "tags.json"), {})
            metrics = _latest_metrics(d)
            ok = True
            for kind, key, op, val in conds:
                src = {"tags": tags, "params": params, "metrics": metrics, "attributes": meta}[kind]
                have = src.get(key)
                if kind == "metrics":
                    try:
                        a, b = float(have), float(val)
                    except (TypeError, ValueError):
                        ok = False
                        break
                    ok = {"=": a == b, "!=": a != b, ">": a > b, "<": a < b, ">=": a >= b, "<=": a <= b}[op]
                else:
                    ok = (str(have) == val) if op == "=" else (str(have) != val) if op == "!=" else False
                if not ok:
                    break
            if not ok:
                continue
            row = {"run_id": rid, "experiment_id": eid, "status": meta.get("status"),
                   "start_time": meta.get("start_time"), "end_time": meta.get("end_time"),
                   "artifact_uri": os.path.join(d, "artifacts")}
            row.update({f"metrics.{k}": v for k, v in metrics.items()})
            row.update({f"params.{k}": v for k, v in params.items()})
            row.update({f"tags.{k}": v for k, v in tags.items()})
            rows.append(row)
    df = pd.DataFrame(rows)
    if df.empty:
        return pd.DataFrame(columns=["run_id", "experiment_id", "status", "start_time", "end_time", "artifact_uri"])
    for ob in reversed(order_by or ["attributes.start_time DESC"]):
        parts = ob.split()
        colname = parts[0].replace("attributes.", "").replace("`", "")
        asc = not (len(parts) > 1 and parts[1].upper() == "DESC")
        if colname not in df.columns:
            df[colname] = float("nan")
        df = df.sort_values(colname, ascending=asc, na_position="last", kind="stable")
    return df.head(max_results).reset_index(drop=True)


# ------------------------------------------------------------------------------------------------ artifact URIs
def resolve_uri(uri: str) -> str:
    """`runs:/<run_id>/<path>`, `models:/<name>/<stage|version>` or a plain path -> local filesystem path."""
    if uri.startswith("runs:/"):
        rid, _, rel = uri[len("runs:/"):].partition("/")
        return os.path.join(_find_run_dir(rid), "artifacts", rel)
    if uri.startswith("models:/"):
        name, _, which = uri[len("models:/"):].partition("/")
        meta = _read_json(os.path.join(get_tracking_uri(), "models", name, "meta.json"), None)
        if meta is None:
            raise KeyError(f"registered model {name!r} not found")
        versions = meta["versions"]
        if which.isdigit():
            cand = [v for v in versions if v["version"] == int(which)]
        else:
            cand = [v for v in versions if v["current_stage"].lower() == which.lower()]
            cand = sorted(cand, key=lambda v: v["version"])[-1:]
        if not cand:
            raise KeyError(f"no version of {name!r} matches {which!r}")
        return resolve_uri(cand[0]["source"])
    return uri[5:] if uri.startswith("file:") else uri


# ------------------------------------------------------------------------------------------------ registry
def register_model(model_uri: str, name: str) -> SimpleNamespace:
    """`mlflow.register_model('runs:/<id>/model', name)` -> ModelVersion (reference P2/01:282-285)."""
    resolve_uri(model_uri)  # must exist
    with _lock:
        d = os.path.join(get_tracking_uri(), "models", name)
        os.makedirs(d, exist_ok=True)
        meta = _read_json(os.path.join(d, "meta.json"), {"name": name, "versions": []})
        v = {"version": len(meta["versions"]) + 1, "source": model_uri, "current_stage": "None",
             "creation_time": time.time()}
        meta["versions"].append(v)
        _write_json(os.path.join(d, "meta.json"), meta)
    return SimpleNamespace(name=name, **v)


class MlflowClient:
    def transition_model_version_stage(self, name: str, version, stage: str,
                                       archive_existing_versions: bool = False) -> SimpleNamespace:
        """None -> Staging -> Production -> Archived (reference P2/01:288-293)."""
        if stage not in ("None", "Staging", "Production", "Archived"):
            raise ValueError(f"invalid stage {stage!r}")
        with _lock:
            p = os.path.join(get_tracking_uri(), "models", name, "meta.json")
            meta = _read_json(p, None)
            if meta is None:
                raise KeyError(name)
            hit = None
            for v in meta["versions"]:
                if v["version"] == int(version):
                    hit = v
                elif archive_existing_versions and v["current_stage"] == stage:
                    v["current_stage"] = "Archived"
            if hit is None:
                raise KeyError(f"{name} has no version {version}")
            hit["current_stage"] = stage
            _write_json(p, meta)
        return SimpleNamespace(name=name, **hit)

    def get_latest_versions(self, name: str, stages: Optional[List[str]] = None) -> List[SimpleNamespace]:
        meta = _read_json(os.path.join(get_tracking_uri(), "models", name, "meta.json"), {"versions": []})
        out = {}
        for v in meta["versions"]:
            if stages is None or v["current_stage"] in stages:
                out[v["current_stage"]] = v
        return [SimpleNamespace(name=name, **v) for v in out.values()]

    def get_run(self, run_id: str) -> Run:
        return get_run(run_id)

    def get_metric_history(self, run_id: str, key: str) -> List[SimpleNamespace]:
        """MLflow's `client.get_metric_history`: Metric-like records (key, value, step, timestamp) in logging order."""
        out = []
        try:
            with open(os.path.join(_find_run_dir(run_id), "metrics.jsonl")) as f:
                for line in f:
                    r = json.loads(line)
                    if r["key"] == key:
                        out.append(SimpleNamespace(key=key, value=r["value"], step=r.get("step"), timestamp=r.get("ts")))
        except FileNotFoundError:
            pass
        return out

    def search_runs(self, experiment_ids, filter_string="", order_by=None):
        return search_runs(experiment_ids, filter_string, order_by)


# ------------------------------------------------------------------------------------------------ autolog
def autolog(disable: bool = False, log_models: bool = True) -> None:
    """`mlflow.autolog()` / `mlflow.tensorflow.autolog()` (reference P1/02:195, P2/01:135): every `Trainer.fit`
    logs its parameters, per-epoch metrics and (rank 0) the final model under artifact path `model`."""
    global _autolog_enabled
    _autolog_enabled = not disable
    from ..train import trainer as _tr
    from ._autolog import install

    install(_tr, enabled=_autolog_enabled, log_models=log_models)


from . import models as keras  # noqa: E402  (mlflow.keras.log_model / load_model spelling)
from . import models  # noqa: E402

tensorflow = SimpleNamespace(autolog=autolog)
mlflow = SimpleNamespace(set_tracking_uri=set_tracking_uri)  # reference typo `mlflow.mlflow.set_tracking_uri` (Q9)

__all__ = ["set_tracking_uri", "get_tracking_uri", "set_experiment", "start_run", "active_run", "end_run", "get_run",
           "log_param", "log_params", "log_metric", "log_metrics", "log_dict", "log_text", "log_artifact", "log_artifacts",
           "set_tag", "set_tags", "get_experiment_by_name",
           "search_runs", "register_model", "MlflowClient", "autolog", "resolve_uri", "keras", "models",
           "metric_history", "get_artifact_uri"]
