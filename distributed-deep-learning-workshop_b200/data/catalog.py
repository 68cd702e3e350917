"""Catalog: databases of named, versioned tables on the local filesystem (the reference's Delta Lake usage,
C3-C8).  A table is a directory of parquet part files (one per partition) plus a `_log/` of JSON commits
(version, mode, files, rows, schema) - `mode('overwrite')` writes a new version and retires the old files,
reads always resolve the latest commit."""
from __future__ import annotations

import json
import os
import re
import shutil
import time
import uuid
from typing import Dict, List, Optional

import pyarrow as pa
import pyarrow.parquet as pq

from .table import Table

_default: Optional["Catalog"] = None


class Conf:
    def __init__(self):
        self._d: Dict[str, str] = {"spark.sql.parquet.compression.codec": "snappy"}

    def set(self, k: str, v) -> None:
        self._d[k] = str(v)

    def get(self, k: str, default=None):
        return self._d.get(k, default)


class Catalog:
    def __init__(self, root: str):
        self.root = os.path.abspath(root)
        os.makedirs(self.root, exist_ok=True)
        self.conf = Conf()
        self.current_database = "default"

    # ---------------------------------------------------------------- databases
    def _db_dir(self, db: str) -> str:
        return os.path.join(self.root, db + ".db")

    def create_database(self, name: str, if_not_exists: bool = True) -> None:
        d = self._db_dir(name)
        if os.path.exists(d) and not if_not_exists:
            raise FileExistsError(f"database {name} already exists")
        os.makedirs(d, exist_ok=True)

    def drop_database(self, name: str, if_exists: bool = True, cascade: bool = False) -> None:
        d = self._db_dir(name)
        if not os.path.exists(d):
            if if_exists:
                return
            raise FileNotFoundError(f"database {name} not found")
        if os.listdir(d) and not cascade:
            raise RuntimeError(f"database {name} is not empty; use cascade")
        shutil.rmtree(d)

    def list_databases(self) -> List[str]:
        return sorted(d[:-3] for d in os.listdir(self.root) if d.endswith(".db"))

    def list_tables(self, db: Optional[str] = None) -> List[str]:
        d = self._db_dir(db or self.current_database)
        return sorted(os.listdir(d)) if os.path.exists(d) else []

    # ---------------------------------------------------------------- tables
    def _split(self, name: str):
        if "." in name:
            db, t = name.split(".", 1)
        else:
            db, t = self.current_database, name
        return db, t

    def _table_dir(self, name: str) -> str:
        db, t = self._split(name)
        return os.path.join(self._db_dir(db), t)

    def _latest_commit(self, tdir: str) -> Optional[dict]:
        log = os.path.join(tdir, "_log")
        if not os.path.isdir(log):
            return None
        versions = sorted(f for f in os.listdir(log) if f.endswith(".json"))
        if not versions:
            return None
        with open(os.path.join(log, versions[-1])) as f:
            return json.load(f)

    def write_table(self, name: str, table: Table, mode: str = "errorifexists", options: Optional[dict] = None) -> None:
        db, _ = self._split(name)
        if not os.path.isdir(self._db_dir(db)):
            raise FileNotFoundError(f"database {db} does not exist (CREATE DATABASE first)")
        tdir = self._table_dir(name)
        prev = self._latest_commit(tdir)
        if prev is not None and mode in ("errorifexists", "error"):
            raise FileExistsError(f"table {name} already exists")
        if prev is not None and mode == "ignore":
            return
        os.makedirs(os.path.join(tdir, "_log"), exist_ok=True)
        codec = (options or {}).get("compression") or self.conf.get("spark.sql.parquet.compression.codec")
        codec = "none" if codec in ("uncompressed", "none", None) else codec
        version = (prev["version"] + 1) if prev else 0
        files = []
        for i, part in enumerate(table.partitions()):
            fn = f"part-{i:05d}-{uuid.uuid4().hex[:8]}.parquet"
            pq.write_table(part, os.path.join(tdir, fn), compression=codec)
            files.append(fn)
        if mode == "append" and prev:
            files = prev["files"] + files
            rows = prev["num_rows"] + table.count()
        else:
            rows = table.count()
        commit = {"version": version, "mode": mode, "files": files, "num_rows": rows, "timestamp": time.time(),
                  "schema": [(f.name, str(f.type)) for f in table.schema], "compression": codec}
        tmp = os.path.join(tdir, "_log", f".{version:020d}.json.tmp")
        with open(tmp, "w") as f:
            json.dump(commit, f)
        os.replace(tmp, os.path.join(tdir, "_log", f"{version:020d}.json"))
        if prev and mode == "overwrite":
            for fn in prev["files"]:
                if fn not in files:
                    try:
                        os.remove(os.path.join(tdir, fn))
                    except FileNotFoundError:
                        pass

    def table(self, name: str) -> Table:
        tdir = self._table_dir(name)
        commit = self._latest_commit(tdir)
        if commit is None:
            raise FileNotFoundError(f"table {name} not found")
        parts = [pq.read_table(os.path.join(tdir, fn)) for fn in commit["files"]]
        t = pa.concat_tables(parts) if parts else pa.table({})
        return Table(t, max(1, len(parts)), self)

    def scan(self, name: str):
        """The table as a lazily read `ScanTable` (one fragment per parquet row group): nothing is loaded here; readers
        such as `pyfunc.spark_udf` workers or `loader.make_converter` open the row groups they own."""
        from .scan import scan_parquet_files

        tdir = self._table_dir(name)
        commit = self._latest_commit(tdir)
        if commit is None:
            raise FileNotFoundError(f"table {name} not found")
        return scan_parquet_files([os.path.join(tdir, fn) for fn in commit["files"]], self)

    def table_history(self, name: str) -> List[dict]:
        log = os.path.join(self._table_dir(name), "_log")
        out = []
        for fn in sorted(os.listdir(log)):
            if fn.endswith(".json"):
                with open(os.path.join(log, fn)) as f:
                    out.append(json.load(f))
        return out

    def drop_table(self, name: str, if_exists: bool = True) -> None:
        tdir = self._table_dir(name)
        if os.path.isdir(tdir):
            shutil.rmtree(tdir)
        elif not if_exists:
            raise FileNotFoundError(name)

    # ---------------------------------------------------------------- the handful of SQL statements the notebooks run
    def sql(self, statement: str) -> Optional[Table]:
        s = statement.strip().rstrip(";")
        m = re.match(r"(?i)^drop\s+database\s+(if\s+exists\s+)?(\w+)(\s+cascade)?$", s)
        if m:
            self.drop_database(m.group(2), if_exists=bool(m.group(1)), cascade=bool(m.group(3)))
            return None
        m = re.match(r"(?i)^create\s+database\s+(if\s+not\s+exists\s+)?(\w+)$", s)
        if m:
            self.create_database(m.group(2), if_not_exists=bool(m.group(1)))
            return None
        m = re.match(r"(?i)^use\s+(\w+)$", s)
        if m:
            self.current_database = m.group(1)
            return None
        m = re.match(r"(?i)^select\s+\*\s+from\s+([\w.]+)(\s+limit\s+(\d+))?$", s)
        if m:
            t = self.table(m.group(1))
            return t.limit(int(m.group(3))) if m.group(3) else t
        m = re.match(r"(?i)^drop\s+table\s+(if\s+exists\s+)?([\w.]+)$", s)
        if m:
            self.drop_table(m.group(2), if_exists=bool(m.group(1)))
            return None
        raise ValueError(f"unsupported SQL statement: {statement!r}")


def default_catalog() -> Catalog:
    global _default
    if _default is None:
        _default = Catalog(os.environ.get("B200DDL_WAREHOUSE", os.path.join(os.getcwd(), "b200ddl_warehouse")))
    return _default


def set_default_catalog(cat: Catalog) -> None:
    global _default
    _default = cat
