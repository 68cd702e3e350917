"""Table: the slice of the Spark DataFrame API the reference's ETL and inference cells use (SURVEY.md L1, C2-C10,
C34), on Apache Arrow.

Verbs (Spark spelling kept as aliases): select, withColumn/with_column, sample, randomSplit/random_split, distinct,
count, collect, limit, repartition, toPandas/to_pandas, display/show, write...saveAsTable.  Column expressions are
vectorised "pandas UDFs" evaluated partition by partition (optionally on a thread pool - the reference's
data-parallel ETL).
"""
from __future__ import annotations

import concurrent.futures as cf
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import pandas as pd
import pyarrow as pa


class Row(dict):
    """collect() element: attribute + item access like pyspark.sql.Row."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


@dataclass
class Column:
    name: str


def col(name: str) -> Column:
    return Column(name)


_ARROW_TYPES = {"string": pa.string(), "int": pa.int32(), "integer": pa.int32(), "long": pa.int64(),
                "bigint": pa.int64(), "float": pa.float32(), "double": pa.float64(), "binary": pa.binary(),
                "boolean": pa.bool_()}


class UDFExpr:
    def __init__(self, udf: "PandasUDF", args: Sequence[str]):
        self.udf, self.args = udf, list(args)


class PandasUDF:
    """`@pandas_udf("string")` - a function pd.Series... -> pd.Series applied per Arrow batch."""

    def __init__(self, fn: Callable, return_type: str = "string"):
        self.fn = fn
        self.return_type = return_type

    def __call__(self, *cols: Union[str, Column]) -> UDFExpr:
        return UDFExpr(self, [c.name if isinstance(c, Column) else c for c in cols])

    def evaluate(self, table: "Table", args: Sequence[str], parallelism: int = 1) -> pa.ChunkedArray:
        parts = table.partitions()

        def run(part: pa.Table) -> pa.Array:
            series = [part.column(a).to_pandas() for a in args]
            out = self.fn(*series)
            typ = _ARROW_TYPES.get(self.return_type)
            if isinstance(out, (pd.Series, np.ndarray, list)):
                return pa.array(np.asarray(out, dtype=object) if typ in (pa.string(), pa.binary()) else np.asarray(out),
                                type=typ, from_pandas=True)
            raise TypeError("pandas UDF must return a Series / ndarray / list")

        if parallelism > 1 and len(parts) > 1:
            with cf.ThreadPoolExecutor(parallelism) as ex:
                arrays = list(ex.map(run, parts))
        else:
            arrays = [run(p) for p in parts]
        return pa.chunked_array(arrays) if arrays else pa.chunked_array([], type=_ARROW_TYPES.get(self.return_type))


def pandas_udf(return_type: Union[str, Callable] = "string"):
    """Decorator: `@pandas_udf("string")` / `@pandas_udf("int")` (reference P1/01:125,187)."""
    if callable(return_type):
        return PandasUDF(return_type, "string")

    def deco(fn):
        return PandasUDF(fn, return_type)

    return deco


class Table:
    def __init__(self, arrow: pa.Table, num_partitions: int = 1, catalog=None):
        self._t = arrow
        self.num_partitions = max(1, int(num_partitions))
        self._catalog = catalog

    # ---------------------------------------------------------------- construction
    @staticmethod
    def from_pandas(df: pd.DataFrame, num_partitions: int = 1) -> "Table":
        return Table(pa.Table.from_pandas(df, preserve_index=False), num_partitions)

    @staticmethod
    def from_pydict(d: Dict[str, Any], num_partitions: int = 1) -> "Table":
        return Table(pa.table(d), num_partitions)

    # ---------------------------------------------------------------- introspection
    @property
    def columns(self) -> List[str]:
        return list(self._t.column_names)

    @property
    def schema(self) -> pa.Schema:
        return self._t.schema

    def to_arrow(self) -> pa.Table:
        return self._t

    def count(self) -> int:
        return self._t.num_rows

    def __len__(self) -> int:
        return self._t.num_rows

    def partitions(self) -> List[pa.Table]:
        n = self._t.num_rows
        k = min(self.num_partitions, max(n, 1))
        bounds = [round(i * n / k) for i in range(k + 1)]
        return [self._t.slice(bounds[i], bounds[i + 1] - bounds[i]) for i in range(k)]

    # ---------------------------------------------------------------- transformations
    def select(self, *cols) -> "Table":
        names = list(cols[0]) if len(cols) == 1 and isinstance(cols[0], (list, tuple)) else list(cols)
        names = [c.name if isinstance(c, Column) else c for c in names]
        return Table(self._t.select(names), self.num_partitions, self._catalog)

    def with_column(self, name: str, expr, parallelism: int = 4) -> "Table":
        if isinstance(expr, UDFExpr):
            arr = expr.udf.evaluate(self, expr.args, parallelism)
        elif hasattr(expr, "evaluate_table"):  # pyfunc shard UDF
            arr = expr.evaluate_table(self)
        elif isinstance(expr, Column):
            arr = self._t.column(expr.name)
        else:
            arr = pa.array(expr)
        t = self._t
        if name in t.column_names:
            t = t.set_column(t.column_names.index(name), name, arr)
        else:
            t = t.append_column(name, arr)
        return Table(t, self.num_partitions, self._catalog)

    withColumn = with_column

    def sample(self, fraction: float, seed: Optional[int] = None, withReplacement: bool = False) -> "Table":
        """Bernoulli sample like Spark's `.sample(fraction=0.5)` (reference P1/01:65)."""
        rng = np.random.default_rng(seed)
        mask = rng.random(self._t.num_rows) < fraction
        return Table(self._t.filter(pa.array(mask)), self.num_partitions, self._catalog)

    def random_split(self, weights: Sequence[float], seed: Optional[int] = None) -> List["Table"]:
        """`randomSplit([0.9, 0.1], seed=42)` (reference P1/01:162)."""
        w = np.asarray(weights, dtype=np.float64)
        edges = np.cumsum(w / w.sum())
        u = np.random.default_rng(seed).random(self._t.num_rows)
        which = np.searchsorted(edges, u, side="right").clip(0, len(w) - 1)
        return [Table(self._t.filter(pa.array(which == i)), self.num_partitions, self._catalog) for i in range(len(w))]

    randomSplit = random_split

    def distinct(self) -> "Table":
        df = self._t.to_pandas().drop_duplicates()
        return Table(pa.Table.from_pandas(df, preserve_index=False), self.num_partitions, self._catalog)

    def limit(self, n: int) -> "Table":
        return Table(self._t.slice(0, n), self.num_partitions, self._catalog)

    def repartition(self, n: int) -> "Table":
        return Table(self._t, n, self._catalog)

    def filter_mask(self, mask: np.ndarray) -> "Table":
        return Table(self._t.filter(pa.array(mask)), self.num_partitions, self._catalog)

    def order_by(self, name: str, ascending: bool = True) -> "Table":
        return Table(self._t.sort_by([(name, "ascending" if ascending else "descending")]), self.num_partitions,
                     self._catalog)

    # ---------------------------------------------------------------- actions
    def collect(self) -> List[Row]:
        return [Row(r) for r in self._t.to_pylist()]

    def to_pandas(self) -> pd.DataFrame:
        return self._t.to_pandas()

    toPandas = to_pandas

    def show(self, n: int = 20) -> None:
        df = self._t.slice(0, n).to_pandas()
        for c in df.columns:
            if df[c].dtype == object and len(df) and isinstance(df[c].iloc[0], (bytes, bytearray)):
                df[c] = df[c].map(lambda b: f"<{len(b)} bytes>")
        print(df.to_string(index=False))

    display = show

    # ---------------------------------------------------------------- writing
    @property
    def write(self) -> "TableWriter":
        return TableWriter(self)


class TableWriter:
    """`df.write.format('delta').mode('overwrite').saveAsTable(name)` (reference P1/01:95,136,216-222)."""

    def __init__(self, table: Table):
        self.table = table
        self._mode = "errorifexists"
        self._format = "delta"
        self._options: Dict[str, str] = {}

    def format(self, fmt: str) -> "TableWriter":
        self._format = fmt
        return self

    def mode(self, m: str) -> "TableWriter":
        self._mode = m
        return self

    def option(self, k: str, v) -> "TableWriter":
        self._options[k] = v
        return self

    def save_as_table(self, name: str, catalog=None) -> None:
        cat = catalog or self.table._catalog
        if cat is None:
            from .catalog import default_catalog

            cat = default_catalog()
        cat.write_table(name, self.table, mode=self._mode, options=self._options)

    saveAsTable = save_as_table
