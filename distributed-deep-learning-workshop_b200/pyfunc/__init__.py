"""Self-contained inference artefacts + sharded batch inference - the `mlflow.pyfunc` surface of the reference
(SURVEY.md L8, C30-C34)::

    class FlowerPyFunc(pyfunc.PythonModel):
        def load_context(self, context): ...            # context.artifacts[name] -> local path
        def predict(self, context, model_input): ...    # pd.Series[bytes] -> np.array[str]

    pyfunc.log_model('pyfunc_model', python_model=FlowerPyFunc(),
                     artifacts={'img_params_dict_path': 'runs:/<id>/img_params_dict.json',
                                'keras_model_path': 'runs:/<id>/model'})               # reference P2/03:354-363
    loaded = pyfunc.load_model('runs:/<id>/pyfunc_model'); loaded.predict(pdf['content'])   # P2/03:446-448
    udf = pyfunc.shard_udf(model_uri, result_type='string')                            # mlflow.pyfunc.spark_udf
    table.with_column('prediction', udf('content'))                                    # P2/03:466-472

`shard_udf` executes as a map over shards of the table: with several GPUs each shard is scored by its own worker
process pinned to one GPU (model loaded once per worker, like a Spark executor's python worker); otherwise the
shards are scored in-process.  Inputs reach `predict` as real `bytes` (no stringified values, unlike the
reference's `ast.literal_eval` workaround at P2/03:228-229, which we also tolerate).
"""
from __future__ import annotations

import json
import os
import shutil
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
import pyarrow as pa


class PythonModel:
    """Base class: override `load_context` (optional) and `predict`."""

    def load_context(self, context) -> None:  # noqa: D401
        pass

    def predict(self, context, model_input):
        raise NotImplementedError


class PythonModelContext(SimpleNamespace):
    """`context.artifacts` maps the names given to `log_model(artifacts=...)` to local paths."""


class PyFuncModel:
    def __init__(self, python_model: PythonModel, context: PythonModelContext, path: str):
        self._impl, self._ctx, self.path = python_model, context, path

    def predict(self, data):
        return self._impl.predict(self._ctx, data)

    def unwrap_python_model(self) -> PythonModel:
        return self._impl


def save_model(path: str, python_model: PythonModel, artifacts: Optional[Dict[str, str]] = None) -> None:
    import cloudpickle

    from .. import tracking

    os.makedirs(os.path.join(path, "artifacts"), exist_ok=True)
    saved = {}
    for name, uri in (artifacts or {}).items():
        src = tracking.resolve_uri(uri)
        if not os.path.exists(src):
            raise FileNotFoundError(f"artifact {name!r}: {uri} -> {src} does not exist")
        dst = os.path.join(path, "artifacts", os.path.basename(src.rstrip("/")))
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True)
        else:
            shutil.copy2(src, dst)
        saved[name] = os.path.relpath(dst, path)
    with open(os.path.join(path, "python_model.pkl"), "wb") as f:
        cloudpickle.dump(python_model, f)
    with open(os.path.join(path, "MLmodel.json"), "w") as f:
        json.dump({"flavor": "b200ddl.pyfunc", "artifacts": saved}, f)


def log_model(artifact_path: str, python_model: PythonModel, artifacts: Optional[Dict[str, str]] = None, **_ignored) -> str:
    from .. import tracking

    save_model(tracking.get_artifact_uri(artifact_path), python_model, artifacts)
    return f"runs:/{tracking.active_run().info.run_id}/{artifact_path}"


def load_model(model_uri: str) -> PyFuncModel:
    import cloudpickle

    from .. import tracking

    path = tracking.resolve_uri(model_uri)
    with open(os.path.join(path, "MLmodel.json")) as f:
        meta = json.load(f)
    with open(os.path.join(path, "python_model.pkl"), "rb") as f:
        impl = cloudpickle.load(f)
    ctx = PythonModelContext(artifacts={k: os.path.join(path, v) for k, v in meta.get("artifacts", {}).items()})
    impl.load_context(ctx)
    return PyFuncModel(impl, ctx, path)


# ------------------------------------------------------------------------------------------------ sharded UDF
_RESULT_TYPES = {"string": pa.string(), "int": pa.int32(), "long": pa.int64(), "double": pa.float64(),
                 "float": pa.float32()}


def _score_shard(args):
    """Worker entry (spawned process): pin a GPU, load the model once, score the shard."""
    model_uri, tracking_uri, shard_index, gpu, values = args
    if gpu is not None and gpu >= 0:
        import torch

        if torch.cuda.is_available():
            torch.cuda.set_device(gpu % torch.cuda.device_count())
    from .. import tracking

    tracking.set_tracking_uri(tracking_uri)
    model = load_model(model_uri)
    out = model.predict(pd.Series(values))
    return shard_index, list(np.asarray(out).tolist())


class ShardUDFExpr:
    def __init__(self, udf: "ShardUDF", column: str):
        self.udf, self.column = udf, column

    def evaluate_table(self, table) -> pa.ChunkedArray:
        return self.udf.evaluate(table, self.column)


class ShardUDF:
    def __init__(self, model_uri: str, result_type: str = "string", num_workers: Optional[int] = None,
                 batch_rows: int = 1024):
        self.model_uri = model_uri
        self.result_type = result_type
        self.num_workers = num_workers
        self.batch_rows = batch_rows
        self._local: Optional[PyFuncModel] = None
        self.stats: Dict[str, Any] = {}

    def __call__(self, column) -> ShardUDFExpr:
        return ShardUDFExpr(self, column.name if hasattr(column, "name") else column)

    def _workers(self) -> int:
        if self.num_workers is not None:
            return max(1, self.num_workers)
        try:
            import torch

            return max(1, torch.cuda.device_count()) if torch.cuda.is_available() else 1
        except Exception:
            return 1

    def evaluate(self, table, column: str) -> pa.ChunkedArray:
        import time

        from .. import tracking

        typ = _RESULT_TYPES.get(self.result_type, pa.string())
        values = table.to_arrow().column(column).to_pylist()
        n = len(values)
        workers = min(self._workers(), max(1, n))
        t0 = time.time()
        if workers <= 1:
            if self._local is None:
                self._local = load_model(self.model_uri)
            outs: List[Any] = []
            for i in range(0, n, self.batch_rows):  # pandas-UDF style: Arrow batch -> Series -> predict
                outs.extend(np.asarray(self._local.predict(pd.Series(values[i:i + self.batch_rows]))).tolist())
        else:
            import multiprocessing as mp

            bounds = [round(i * n / workers) for i in range(workers + 1)]
            jobs = [(self.model_uri, tracking.get_tracking_uri(), i, i, values[bounds[i]:bounds[i + 1]])
                    for i in range(workers) if bounds[i + 1] > bounds[i]]
            ctx = mp.get_context("spawn")
            with ctx.Pool(len(jobs)) as pool:
                res = dict(pool.map(_score_shard, jobs))
            outs = [v for i in sorted(res) for v in res[i]]
        self.stats = {"rows": n, "workers": workers, "seconds": time.time() - t0,
                      "rows_per_sec": n / max(time.time() - t0, 1e-9)}
        return pa.chunked_array([pa.array(outs, type=typ)])


def shard_udf(model_uri: str, result_type: str = "string", num_workers: Optional[int] = None) -> ShardUDF:
    return ShardUDF(model_uri, result_type, num_workers)


def spark_udf(spark, model_uri: str, result_type: str = "string") -> ShardUDF:
    """Reference spelling: `mlflow.pyfunc.spark_udf(spark, model_uri, result_type='string')` (P2/03:466)."""
    return ShardUDF(model_uri, result_type)


__all__ = ["PythonModel", "PythonModelContext", "PyFuncModel", "log_model", "save_model", "load_model", "shard_udf",
           "spark_udf", "ShardUDF"]
