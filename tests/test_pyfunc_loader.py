"""pyfunc packaging + batch inference (reference P2/03) and the converter/ring loader (reference Petastorm usage)."""
import io
import json

import numpy as np
import pandas as pd
import pytest
import torch

from b200ddl import Session, optim, pyfunc, tracking
from b200ddl.data import synthetic_images, pandas_udf, col
from b200ddl.loader import make_converter
from b200ddl.models import CLASSES, build_model, decode_image
from b200ddl.train import Trainer

IMG = 32


class FlowerPyFunc(pyfunc.PythonModel):                       # reference P2/03:157-234
    def load_context(self, context):
        with open(context.artifacts["img_params_dict_path"]) as f:
            d = json.load(f)
        self.img_height, self.img_width = d["img_height"], d["img_width"]
        self.model = tracking.keras.load_model(context.artifacts["keras_model_path"])

    def preprocess(self, img_bytes):
        return decode_image(img_bytes, (self.img_height, self.img_width))

    def predict(self, context, model_input: pd.Series) -> np.ndarray:
        arr = np.stack([self.preprocess(b) for b in model_input])
        logits = self.model.predict(arr, batch_size=16)
        return np.take(CLASSES, np.argmax(logits, axis=1))


@pytest.fixture()
def session(tmp_path):
    s = Session(user="me@example.com", root=str(tmp_path))
    tracking.set_experiment("pyfunc")
    return s


def _tables(session, n=48):
    raw = synthetic_images(n, size=(IMG, IMG), jpeg=True, seed=2)

    @pandas_udf("string")
    def label(path):
        return path.map(lambda p: p.split("/")[-2])

    @pandas_udf("int")
    def label_idx(lab):
        return lab.map(lambda l: CLASSES.index(l))

    return raw.withColumn("label", label(col("path"))).withColumn("label_idx", label_idx(col("label")))


def test_converter_len_sharding_and_infinite_epochs(session):
    t = _tables(session, 40).select(["content", "label_idx"])
    conv = make_converter(t, session.cache_dir)
    assert len(conv) == 40                                                         # reference P1/03:143
    seen = []
    for shard in range(2):
        with conv.make_dataset(batch_size=4, cur_shard=shard, shard_count=2, workers_count=1, image_size=(IMG, IMG),
                               device="cpu") as ds:                                # reference P1/03:332-337
            labs = []
            for _ in range(7):                                                     # 20 rows / 4 = 5 batches per epoch -> wraps
                x, y = next(ds)
                assert x.shape == (4, IMG, IMG, 3) and x.dtype == torch.uint8 and y.dtype == torch.int64
                labs.append(y.clone())
            seen.append(torch.cat(labs)[:20])
    full = t.to_pandas()["label_idx"].to_numpy()
    assert sorted(torch.cat(seen).tolist()) == sorted(full.tolist())                # disjoint shards cover the table
    with pytest.raises(ValueError):
        conv.make_dataset(4, cur_shard=0)
    conv.delete()
    import os
    assert not os.path.exists(conv.cache_dir)                                      # reference P1/03:425-426


def test_finite_epochs_end_the_iteration(session):
    """`num_epochs=k` must END: the consumer gets every full batch the rows allow, then StopIteration (it used to block
    forever once the decode workers had run out of rows) - also through `Trainer.evaluate(ds)` without a step count."""
    import threading

    t = _tables(session, 37).select(["content", "label_idx"])
    conv = make_converter(t, session.cache_dir)
    full = sorted(t.to_pandas()["label_idx"].tolist())
    for epochs, workers, batches in ((1, 2, 4), (2, 3, 9), (1, 1, 4)):      # 37 rows, batch 8: 4 per epoch, 9 over two
        got = []

        def consume():
            with conv.make_dataset(batch_size=8, num_epochs=epochs, workers_count=workers, image_size=(IMG, IMG),
                                   device="cpu") as ds:
                for x, y in ds:
                    assert x.shape == (8, IMG, IMG, 3)
                    got.append(y.tolist())

        th = threading.Thread(target=consume, daemon=True)
        th.start()
        th.join(timeout=60)
        assert not th.is_alive(), f"iteration over a finite dataset did not end (epochs={epochs}, workers={workers})"
        assert len(got) == batches
        if epochs == 1 and workers == 1:                                      # one worker: file order, tail of 5 rows dropped
            assert [l for b in got for l in b] == t.to_pandas()["label_idx"].tolist()[:32]
        from collections import Counter
        assert not Counter(l for b in got for l in b) - Counter(full * epochs)   # only rows of the table, each at most `epochs` times
    # len(ds) = full batches per epoch of the shard: 37 rows -> shards of 19 and 18 rows -> 2 and 2 batches of 8
    with conv.make_dataset(batch_size=8, cur_shard=0, shard_count=2, num_epochs=1, workers_count=1, image_size=(IMG, IMG),
                           device="cpu") as d0, \
         conv.make_dataset(batch_size=8, cur_shard=1, shard_count=2, num_epochs=1, workers_count=1, image_size=(IMG, IMG),
                           device="cpu") as d1, \
         conv.make_dataset(batch_size=8, num_epochs=None, workers_count=1, image_size=(IMG, IMG), device="cpu") as dall:
        assert (len(d0), len(d1), len(dall)) == (2, 2, 4)
        assert sum(1 for _ in d0) == 2 and sum(1 for _ in d1) == 2
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * IMG * IMG, len(CLASSES)))
    tr = Trainer(model, device="cpu").compile(optimizer=optim.SGD(0.0))
    with conv.make_dataset(batch_size=8, num_epochs=None, workers_count=2, image_size=(IMG, IMG), device="cpu") as ds:
        h = tr.fit(ds, epochs=2, verbose=0)                                   # steps_per_epoch defaults to len(ds)
        assert len(h.history["loss"]) == 2 and tr.steps_per_epoch == 4
    done = []

    def evaluate():
        with conv.make_dataset(batch_size=8, num_epochs=1, workers_count=2, image_size=(IMG, IMG), device="cpu") as ds:
            done.append(tr.evaluate(ds))

    th = threading.Thread(target=evaluate, daemon=True)
    th.start()
    th.join(timeout=60)
    assert done and np.isfinite(done[0][0])
    conv.delete()


def test_decoded_pixels_match_pil(session):
    t = _tables(session, 4).select(["content", "label_idx"])
    conv = make_converter(t, session.cache_dir)
    with conv.make_dataset(batch_size=4, workers_count=1, image_size=(IMG, IMG), device="cpu", num_epochs=1) as ds:
        x, y = next(ds)
    from PIL import Image

    ref = np.stack([np.asarray(Image.open(io.BytesIO(c)).convert("RGB")) for c in t.to_pandas()["content"]])
    assert np.array_equal(x.numpy(), ref)
    conv.delete()


def test_train_log_pyfunc_and_batch_inference(session):
    data = _tables(session, 48)
    conv = make_converter(data.select(["content", "label_idx"]), session.cache_dir)
    tracking.autolog()
    try:
        with tracking.start_run(run_name="pyfunc_model_petastorm") as run:         # reference P2/03:280
            tracking.log_dict({"img_height": IMG, "img_width": IMG}, "img_params_dict.json")
            model = build_model(IMG, IMG, 3, 5, arch="mobilenetv2", freeze_base=False)
            tr = Trainer(model, device="cpu").compile(optimizer=optim.Adam(2e-3))
            with conv.make_dataset(batch_size=8, workers_count=2, image_size=(IMG, IMG), device="cpu") as ds:
                tr.fit(ds, steps_per_epoch=len(conv) // 8, epochs=2, verbose=0)
            rid = run.info.run_id
            pyfunc.log_model("pyfunc_model", python_model=FlowerPyFunc(),
                             artifacts={"img_params_dict_path": f"runs:/{rid}/img_params_dict.json",
                                        "keras_model_path": f"runs:/{rid}/model"})   # reference P2/03:354-363
    finally:
        tracking.autolog(disable=True)
    conv.delete()
    loaded = pyfunc.load_model(f"runs:/{rid}/pyfunc_model")                         # reference P2/03:446
    pdf = data.limit(10).toPandas()
    pred = loaded.predict(pdf["content"])                                           # reference P2/03:448
    assert pred.shape == (10,) and set(pred) <= set(CLASSES)
    udf = pyfunc.spark_udf(None, f"runs:/{rid}/pyfunc_model", result_type="string")  # reference P2/03:466
    udf.num_workers = 1
    out = data.limit(20).withColumn("prediction", udf("content")).select("path", "content", "label", "prediction")
    df = out.to_pandas()
    assert len(df) == 20 and list(df["prediction"][:10]) == list(pred)
    assert udf.stats["rows"] == 20
    # the multi-process branch (one spawned scoring process per GPU on a GPU box): the workers must be able to load the
    # user's PythonModel class, which lives in THIS module, and the shards must come back in row order
    udf3 = pyfunc.spark_udf(None, f"runs:/{rid}/pyfunc_model", result_type="string")
    udf3.num_workers = 3
    df3 = data.limit(20).withColumn("prediction", udf3("content")).to_pandas()
    assert udf3.stats["workers"] == 3 and list(df3["prediction"]) == list(df["prediction"])
    # more workers than rows, and an empty table
    udf9 = pyfunc.shard_udf(f"runs:/{rid}/pyfunc_model", num_workers=9)
    assert list(data.limit(2).withColumn("prediction", udf9("content")).to_pandas()["prediction"]) == list(df["prediction"][:2])
    assert udf9.stats["workers"] == 2
    assert len(data.limit(0).withColumn("prediction", udf9("content")).to_pandas()) == 0


def test_row_group_sharding_reads_each_group_once_and_shuffle_changes_order(session):
    """Shard s owns a contiguous row range and opens only the row groups that intersect it (8 ranks together read the
    cache once, not 8 times); `shuffle=True` permutes row groups and rows with seed + epoch, deterministically."""
    from b200ddl.loader import Converter

    t = _tables(session, 96).select(["content", "label_idx"])
    conv = Converter(t, session.cache_dir, rows_per_group=8)                   # 12 row groups
    labels = t.to_pandas()["label_idx"].tolist()
    seen, groups = [], 0
    for shard in range(4):
        with conv.make_dataset(batch_size=8, cur_shard=shard, shard_count=4, num_epochs=1, workers_count=1,
                               image_size=(IMG, IMG), device="cpu") as ds:
            got = [l for _, y in ds for l in y.tolist()]
            assert got == labels[24 * shard:24 * (shard + 1)]                     # contiguous range, file order
            seen += got
            groups += ds.row_groups_read
    assert seen == labels and groups == 12                                       # every row group opened exactly once

    def epoch_labels(seed, epochs=1):
        with conv.make_dataset(batch_size=8, num_epochs=epochs, workers_count=1, image_size=(IMG, IMG), device="cpu",
                               shuffle=True, seed=seed) as ds:
            return [l for _, y in ds for l in y.tolist()]

    a, b, c = epoch_labels(1), epoch_labels(1), epoch_labels(2)
    assert a == b and sorted(a) == sorted(labels)                                # same seed -> same order, a permutation
    assert a != labels and c != a                                                # not the file order; seed changes it
    two = epoch_labels(1, epochs=2)
    assert two[:96] == a and two[96:] != a and sorted(two[96:]) == sorted(labels)  # epoch 2 is reshuffled
    conv.delete()


class MeanPixel(pyfunc.PythonModel):
    def predict(self, context, model_input: pd.Series) -> np.ndarray:
        from b200ddl.models import decode_batch
        a = decode_batch(model_input, (IMG, IMG))
        return a.reshape(len(a), -1).mean(1).astype(np.float64)


def test_shard_udf_over_lazy_scan_pool_warmup_and_stats(session):
    """BASELINE config 4's shape on CPU: a lazily generated table scored by a persistent worker pool.  Start-up (process
    start, model load, warm-up) is reported apart from the scored seconds; a table with another fragment shape re-warms the
    pool once; results are Arrow arrays built in the workers and come back in row order."""
    from b200ddl.data import synthetic_scan

    with tracking.start_run():
        uri = pyfunc.log_model("mean_pixel", python_model=MeanPixel())
    udf = pyfunc.shard_udf(uri, "double", num_workers=2)
    try:
        t1 = synthetic_scan(300, size=(IMG, IMG), rows_per_fragment=100)
        p1 = t1.withColumn("p", udf("content")).select("p").toPandas()["p"].to_numpy()
        s1 = dict(udf.stats)
        assert s1["workers"] == 2 and s1["fragments"] == 3 and s1["rows"] == 300
        assert s1["startup_seconds"] > 0                                # reported apart from `seconds`
        ref = MeanPixel().predict(None, t1.select("content").toPandas()["content"])
        assert np.allclose(p1, ref)
        pool = udf._pool
        t2 = synthetic_scan(800, size=(IMG, IMG), rows_per_fragment=200)
        p2 = t2.withColumn("p", udf("content")).select("p").toPandas()["p"].to_numpy()
        s2 = dict(udf.stats)
        assert udf._pool is pool and s2["startup_seconds"] > 0          # same processes, re-warmed for the new fragment shape
        p3 = t2.withColumn("p", udf("content")).select("p").toPandas()["p"].to_numpy()
        s3 = dict(udf.stats)
        assert s3["startup_seconds"] == 0.0 and np.array_equal(p2, p3)
        # worker statistics are per job, not cumulative over the pool's life
        assert sum(w["predict_s"] for w in s3["per_worker"]) < 1.5 * sum(w["predict_s"] for w in s2["per_worker"]) + 0.05
        assert all(w["wall_s"] >= w["predict_s"] * 0.5 for w in s3["per_worker"]) and len(s3["per_worker"]) == 2
    finally:
        udf.close()


def test_decode_processes_match_decode_threads_and_surface_worker_death(session):
    """`make_dataset(decode_processes=N)`: the decode runs in N torch-free worker processes whose output lands directly in
    the ring's slots - same batches, bit for bit, as the thread path; a decode process that dies ends the stream with an
    error instead of a hang."""
    data = _tables(session, 64)
    conv = make_converter(data.select(["content", "label_idx"]), session.cache_dir)
    with conv.make_dataset(batch_size=16, num_epochs=1, workers_count=1, image_size=(IMG, IMG), device="cpu") as a, \
         conv.make_dataset(batch_size=16, num_epochs=1, workers_count=1, image_size=(IMG, IMG), device="cpu",
                           decode_processes=3) as b:
        assert b.decode_processes == 3 and a.decode_processes == 0
        n = 0
        for (xa, ya), (xb, yb) in zip(a, b):
            assert torch.equal(xa, xb) and torch.equal(ya, yb)
            n += 1
        assert n == 4
    # two filler threads x two processes each, infinite epochs, shuffled: every batch is complete and the labels are valid
    with conv.make_dataset(batch_size=8, workers_count=2, image_size=(IMG, IMG), device="cpu", shuffle=True, seed=3,
                           decode_processes=4) as ds:
        it = iter(ds)
        for _ in range(12):
            x, y = next(it)
            assert x.shape == (8, IMG, IMG, 3) and int(y.min()) >= 0 and int(y.max()) < 5 and float(x.float().std()) > 1.0
        # kill one decode process: the stream must end with an error, not block
        ds._pools[0][0].kill()
        with pytest.raises(RuntimeError, match="decode pipeline failed"):
            for _ in range(64):
                next(it)
    conv.delete()


def _planner_ns(files, total, shard, count, shuffle=False, seed=0, num_epochs=1):
    """The row planner of the table datasets without a ring / decode threads (what `_GpuDecodeDataset` borrows too)."""
    import types

    from b200ddl.loader import _TableDataset

    ns = types.SimpleNamespace(files=files, total_rows=total, cur_shard=shard, shard_count=count, num_epochs=num_epochs,
                               shuffle=shuffle, seed=seed, row_groups_read=0, batch_size=1)
    ns.row_lo, ns.row_hi = (total * shard) // count, (total * (shard + 1)) // count
    ns._plan = lambda: _TableDataset._plan(ns)
    return ns, _TableDataset._rows(ns)


def test_loader_sharding_properties(tmp_path):
    """Property test: for arbitrary file / row-group layouts and shard counts, the shards are
    disjoint, together cover every row exactly once, differ in size by at most one row, open only row groups they own rows
    of, and `shuffle` yields a seed-determined permutation of exactly the shard's rows that changes from epoch to epoch."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from hypothesis import given, settings, strategies as st

    counter = [0]

    @settings(max_examples=20, deadline=None)
    @given(files=st.lists(st.tuples(st.integers(1, 40), st.integers(1, 17)), min_size=1, max_size=3),
           shards=st.integers(1, 7), seed=st.integers(0, 1000))
    def check(files, shards, seed):
        counter[0] += 1
        paths, base = [], 0
        for k, (rows, rg) in enumerate(files):
            p = tmp_path / f"c{counter[0]}_{k}.parquet"
            ids = list(range(base, base + rows))
            pq.write_table(pa.table({"content": [i.to_bytes(4, "little") for i in ids], "label_idx": ids}), p, row_group_size=rg)
            paths.append((str(p), base))
            base += rows
        total = base
        seen, sizes = [], []
        for s in range(shards):
            ns, rows = _planner_ns(paths, total, s, shards)
            got = [int(labels[j]) for _, labels, j in rows]
            assert got == list(range(ns.row_lo, ns.row_hi))                      # contiguous range, file order
            groups_owned = sum(1 for (fi, rg, skip, cnt) in ns._plan())
            assert ns.row_groups_read == groups_owned                             # no read amplification
            seen += got
            sizes.append(len(got))
            ns1, r1 = _planner_ns(paths, total, s, shards, shuffle=True, seed=seed, num_epochs=2)
            ns2, r2 = _planner_ns(paths, total, s, shards, shuffle=True, seed=seed, num_epochs=2)
            a = [int(l[j]) for _, l, j in r1]
            b = [int(l[j]) for _, l, j in r2]
            assert a == b and len(a) == 2 * len(got)                              # deterministic for a seed, two epochs
            assert sorted(a[:len(got)]) == got and sorted(a[len(got):]) == got    # every epoch = a permutation of the shard
        assert sorted(seen) == list(range(total)) and max(sizes) - min(sizes) <= 1

    check()
