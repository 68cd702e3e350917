"""ScanTable: a table that is a list of *fragments* read (or generated) where they are consumed.

The reference scores `spark.table(...)` with `mlflow.pyfunc.spark_udf` (P2/03:466-472): Spark never brings the image
bytes to the driver - every executor reads its own partitions.  `Catalog.scan(name)` / `synthetic_scan(...)` give the
same property here: a `ScanTable` knows its schema and row counts, `limit / select / count` are metadata operations,
and `with_column(name, shard_udf(col))` hands the *fragment descriptors* (file + row group, or generator spec - a few
hundred bytes each) to the scoring workers, which read their own bytes.  Only the predictions come back.

Anything else (`to_pandas`, `collect`, ...) materialises the selected columns, like `toPandas()` does in Spark.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

from .table import Column, Table


@dataclass(frozen=True)
class Fragment:
    """Picklable description of a slice of rows.

    kind 'parquet': rows [skip, skip+rows) of row group `row_group` of file `path`;
    kind 'ipc':     record batch `row_group` of the Arrow IPC file `path` (memory-mapped, zero copy);
    kind 'synthetic': rows [start, start+rows) of a generated image table (see `synthetic_scan`)."""
    kind: str
    rows: int
    path: str = ""
    row_group: int = 0
    skip: int = 0
    start: int = 0
    spec: Tuple = ()

    def read(self, columns: Optional[Sequence[str]] = None, pinned_alloc=None) -> pa.Table:
        if self.kind == "parquet":
            import pyarrow.parquet as pq

            t = pq.ParquetFile(self.path).read_row_group(self.row_group, columns=list(columns) if columns else None)
            return t.slice(self.skip, self.rows)
        if self.kind == "ipc":
            with pa.memory_map(self.path, "r") as mm:
                b = pa.ipc.open_file(mm).get_batch(self.row_group)
                t = pa.Table.from_batches([b])
            t = t.slice(self.skip, self.rows)
            return t.select(list(columns)) if columns else t
        if self.kind == "synthetic":
            return _synthetic_read(self, columns, pinned_alloc)
        raise ValueError(f"unknown fragment kind {self.kind!r}")


# ------------------------------------------------------------------------------------------------ synthetic images
_POOLS: Dict[Tuple, Tuple[np.ndarray, np.ndarray]] = {}


def _synthetic_pool(spec) -> Tuple[np.ndarray, np.ndarray]:
    """(images uint8 [P, H*W*3], labels int64 [P]) - class-coloured noise, generated once per process."""
    h, w, n_classes, pool, seed = spec
    key = (h, w, n_classes, pool, seed)
    if key not in _POOLS:
        rng = np.random.default_rng(seed)
        labels = rng.integers(0, n_classes, size=pool).astype(np.int64)
        base = rng.uniform(40, 215, size=(n_classes, 3))
        imgs = np.empty((pool, h * w * 3), dtype=np.uint8)
        # a bank of 8 noise tiles, shifted differently per image: 8 x (h*w*3) normals instead of pool x (h*w*3) - building
        # the pool was ~0.4 s of a scoring worker's first fragment
        bank = rng.standard_normal(size=(8, h, w, 3), dtype=np.float32) * np.float32(25.0)
        for i in range(pool):
            img = np.roll(bank[i % 8], ((i * 7) % max(h, 1), (i * 13) % max(w, 1)), axis=(0, 1)) + base[labels[i]].astype(np.float32)
            imgs[i] = np.clip(img, 0, 255).astype(np.uint8).reshape(-1)
        _POOLS[key] = (imgs, labels)
    return _POOLS[key]


def _synthetic_read(f: Fragment, columns, pinned_alloc) -> pa.Table:
    from .sources import FLOWER_CLASSES

    h, w, n_classes, pool, seed = f.spec
    imgs, labels = _synthetic_pool(f.spec)
    # row r of the table is pool image r % pool: a fragment is a handful of LONG contiguous runs of the pool, so filling
    # it is a few large memcpys (measured: np.take of scattered 150 KB rows reaches only ~2-4 GB/s per process, which
    # capped a scoring worker at ~27 k images/s - below what its GPU consumes)
    idx = (np.arange(f.start, f.start + f.rows, dtype=np.int64) % pool)
    want = list(columns) if columns else ["path", "length", "content", "label", "label_idx"]
    cols = {}
    row_bytes = h * w * 3
    if "content" in want:
        nbytes = f.rows * row_bytes
        # the image bytes land in ONE contiguous buffer (pinned when the caller provides an allocator), which becomes the
        # Arrow binary column's data buffer without a copy - and later the source of the H2D copy
        buf = pinned_alloc(nbytes) if pinned_alloc is not None else np.empty(nbytes, dtype=np.uint8)
        view = buf[:nbytes].reshape(f.rows, row_bytes)
        _copy_runs(imgs, int(f.start % pool), view)
        offsets = (np.arange(f.rows + 1, dtype=np.int64) * row_bytes)
        if nbytes < 2 ** 31:
            arr = pa.Array.from_buffers(pa.binary(), f.rows, [None, pa.py_buffer(offsets.astype(np.int32)), pa.py_buffer(view)])
        else:
            arr = pa.Array.from_buffers(pa.large_binary(), f.rows, [None, pa.py_buffer(offsets), pa.py_buffer(view)])
        cols["content"] = arr
    lab = labels[idx]
    names = FLOWER_CLASSES if n_classes == len(FLOWER_CLASSES) else [f"class_{i}" for i in range(n_classes)]
    if "label_idx" in want:
        cols["label_idx"] = pa.array(lab, pa.int64())
    if "label" in want:
        cols["label"] = pa.array(np.take(np.asarray(names, dtype=object), lab), pa.string())
    if "path" in want:
        cols["path"] = pa.array([f"synthetic:/{names[l]}/{f.start + i:09d}.raw" for i, l in enumerate(lab)], pa.string())
    if "length" in want:
        cols["length"] = pa.array(np.full(f.rows, row_bytes, dtype=np.int64))
    return pa.table({k: cols[k] for k in want if k in cols})


def _copy_runs(pool_imgs: np.ndarray, first: int, out: np.ndarray) -> None:
    """out[r] = pool_imgs[(first + r) % P] as contiguous block copies (numpy releases the GIL for them), a few in parallel."""
    import concurrent.futures as cf

    P, n = pool_imgs.shape[0], out.shape[0]
    runs, r, src = [], 0, first
    chunk = max(1, min(P, 64))  # <= 64 images (~10 MB at 224x224x3) per copy keeps several threads busy
    while r < n:
        k = min(chunk, P - src, n - r)
        runs.append((r, src, k))
        r += k
        src = (src + k) % P
    threads = int(os.environ.get("B200DDL_GEN_THREADS", "6"))

    def do(run):
        r0, s0, k = run
        np.copyto(out[r0:r0 + k], pool_imgs[s0:s0 + k])

    if threads > 1 and len(runs) > 1:
        with cf.ThreadPoolExecutor(threads) as ex:
            list(ex.map(do, runs))
    else:
        for run in runs:
            do(run)


def _parallel_take(src: np.ndarray, idx: np.ndarray, out: np.ndarray, threads: int = 4) -> None:
    """out[i] = src[idx[i]] with a few threads (numpy releases the GIL inside take for large copies)."""
    n = len(idx)
    if n < 64 or threads <= 1:
        np.take(src, idx, axis=0, out=out)
        return
    import concurrent.futures as cf

    bounds = [round(i * n / threads) for i in range(threads + 1)]
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: np.take(src, idx[bounds[k]:bounds[k + 1]], axis=0, out=out[bounds[k]:bounds[k + 1]]),
                    range(threads)))


_SYNTH_SCHEMA = pa.schema([("path", pa.string()), ("length", pa.int64()), ("content", pa.binary()),
                           ("label", pa.string()), ("label_idx", pa.int64())])


# ------------------------------------------------------------------------------------------------ the table
class ScanTable(Table):
    """Lazily evaluated table over fragments (+ columns computed by UDFs, kept in memory and aligned by row)."""

    def __init__(self, fragments: List[Fragment], schema: pa.Schema, catalog=None,
                 projection: Optional[List[str]] = None, attached: Optional[Dict[str, pa.ChunkedArray]] = None):
        self._fragments = list(fragments)
        self._schema = schema
        self._catalog = catalog
        self._projection = projection
        self._attached = dict(attached or {})
        self.num_partitions = max(1, len(self._fragments))
        self._cache: Optional[pa.Table] = None

    # -- metadata operations ----------------------------------------------------------------------------------
    def fragments(self) -> List[Fragment]:
        return list(self._fragments)

    @property
    def columns(self) -> List[str]:
        base = self._projection if self._projection is not None else [f.name for f in self._schema]
        return list(base) + [c for c in self._attached if c not in base]

    @property
    def schema(self) -> pa.Schema:
        fields = [self._schema.field(c) if c in self._schema.names else pa.field(c, self._attached[c].type)
                  for c in self.columns]
        return pa.schema(fields)

    def count(self) -> int:
        return sum(f.rows for f in self._fragments)

    __len__ = count

    def _derive(self, **kw) -> "ScanTable":
        args = dict(fragments=self._fragments, schema=self._schema, catalog=self._catalog, projection=self._projection,
                    attached=self._attached)
        args.update(kw)
        return ScanTable(**args)

    def limit(self, n: int) -> "ScanTable":
        out, left = [], int(n)
        for f in self._fragments:
            if left <= 0:
                break
            take = min(f.rows, left)
            out.append(f if take == f.rows else replace(f, rows=take))
            left -= take
        att = {k: v.slice(0, n) for k, v in self._attached.items()}
        return self._derive(fragments=out, attached=att)

    def select(self, *cols) -> "ScanTable":
        names = list(cols[0]) if len(cols) == 1 and isinstance(cols[0], (list, tuple)) else list(cols)
        names = [c.name if isinstance(c, Column) else c for c in names]
        for c in names:
            if c not in self._schema.names and c not in self._attached:
                raise KeyError(f"no column {c!r}; have {self.columns}")
        return self._derive(projection=[c for c in names if c in self._schema.names],
                            attached={c: self._attached[c] for c in names if c in self._attached})

    def repartition(self, n: int) -> "ScanTable":
        return self  # partitioning = fragments

    def with_column(self, name: str, expr, parallelism: int = 4):
        if hasattr(expr, "evaluate_table"):  # pyfunc shard UDF: workers read their own fragments
            arr = expr.evaluate_table(self)
            att = dict(self._attached)
            att[name] = arr if isinstance(arr, pa.ChunkedArray) else pa.chunked_array([arr])
            return self._derive(attached=att)
        return self._materialised().with_column(name, expr, parallelism)

    withColumn = with_column

    # -- materialisation --------------------------------------------------------------------------------------
    def to_arrow(self) -> pa.Table:
        if self._cache is None:
            base = self._projection if self._projection is not None else list(self._schema.names)
            parts = [f.read(base) for f in self._fragments] if base else []
            t = pa.concat_tables(parts) if parts else pa.table({})
            for k, v in self._attached.items():
                t = t.append_column(k, v) if base else pa.table({k: v})
            self._cache = t
        return self._cache

    def _materialised(self) -> Table:
        return Table(self.to_arrow(), self.num_partitions, self._catalog)

    @property
    def _t(self) -> pa.Table:  # every inherited eager verb (sample, distinct, collect, toPandas, ...) works on this
        return self.to_arrow()

    @_t.setter
    def _t(self, v) -> None:
        pass


def synthetic_scan(n: int, size: Tuple[int, int] = (224, 224), num_classes: int = 5, rows_per_fragment: int = 4096,
                   pool_images: int = 256, seed: int = 0) -> ScanTable:
    """`n` JPEG-shaped synthetic images (raw uint8 H*W*3 payloads, class-coloured noise drawn from a pool of
    `pool_images` distinct images) as a lazily generated table: every fragment is produced by whoever reads it."""
    h, w = size
    rows_per_fragment = max(1, min(rows_per_fragment, (2 ** 31 - 1) // (h * w * 3)))
    spec = (h, w, int(num_classes), int(pool_images), int(seed))
    frags = [Fragment("synthetic", min(rows_per_fragment, n - s), start=s, spec=spec)
             for s in range(0, n, rows_per_fragment)]
    return ScanTable(frags, _SYNTH_SCHEMA)


def scan_parquet_files(paths: Sequence[str], catalog=None) -> ScanTable:
    import pyarrow.parquet as pq

    frags: List[Fragment] = []
    schema = None
    for p in paths:
        pf = pq.ParquetFile(p)
        if schema is None:
            schema = pf.schema_arrow
        for rg in range(pf.num_row_groups):
            frags.append(Fragment("parquet", pf.metadata.row_group(rg).num_rows, path=os.path.abspath(p), row_group=rg))
    return ScanTable(frags, schema if schema is not None else pa.schema([]), catalog)
