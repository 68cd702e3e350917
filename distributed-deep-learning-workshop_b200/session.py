"""Session: what `%run ./00_setup` gives every reference notebook (C1; P1/00:3-17) - a per-user namespace
(`user`, `my_name`, `database_name`), where the tracking server lives, and the ambient `spark` / `dbutils`-like
handles (`catalog`, `fs`)."""
from __future__ import annotations

import getpass
import os
import re
import shutil
from typing import List, Optional


class FS:
    """`dbutils.fs.{rm, mkdirs, ls}` on the local filesystem (accepts `file:` / `dbfs:` prefixes)."""

    def __init__(self, root: str):
        self.root = root

    def _p(self, path: str) -> str:
        path = re.sub(r"^(file:|dbfs:)/*", "/", path)
        return path if os.path.isabs(path) and not path.startswith("/dbfs") else os.path.join(self.root, path.lstrip("/"))

    def rm(self, path: str, recurse: bool = False) -> bool:
        p = self._p(path)
        if os.path.isdir(p):
            shutil.rmtree(p) if recurse else os.rmdir(p)
            return True
        if os.path.exists(p):
            os.remove(p)
            return True
        return False

    def mkdirs(self, path: str) -> bool:
        os.makedirs(self._p(path), exist_ok=True)
        return True

    def ls(self, path: str) -> List[str]:
        p = self._p(path)
        return sorted(os.path.join(p, f) for f in os.listdir(p))


class Session:
    def __init__(self, user: Optional[str] = None, root: Optional[str] = None):
        self.user = user or os.environ.get("B200DDL_USER") or f"{getpass.getuser()}@localhost"
        # my_name: local part of the e-mail with '.' -> '_' (reference P1/00:6); single sanitised namespace (Q10)
        self.my_name = re.sub(r"[^0-9a-zA-Z_]", "_", self.user.split("@")[0])
        self.database_name = f"distributed_dl_workshop_{self.my_name}"
        self.root = os.path.abspath(root or os.environ.get("B200DDL_HOME", os.path.join(os.getcwd(), "b200ddl_home")))
        os.makedirs(self.root, exist_ok=True)
        self.tracking_uri = os.path.join(self.root, "mlruns")
        self.cache_dir = os.path.join(self.root, "tmp", f"distributed_dl_workshop_{self.my_name}", "petastorm")
        self.checkpoint_root = os.path.join(self.root, f"distributed_dl_workshop_{self.my_name}", "train_ckpts")
        from .data import Catalog, set_default_catalog
        from . import tracking

        self.catalog = Catalog(os.path.join(self.root, "warehouse"))
        set_default_catalog(self.catalog)
        tracking.set_tracking_uri(self.tracking_uri)
        self.fs = FS(self.root)

    # the reference captures host/token so WORKER processes can reach the tracking server (P1/03:286-288);
    # here the equivalent is the tracking URI, which `Runner` children inherit through the environment.
    @property
    def DATABRICKS_HOST(self) -> str:
        return self.tracking_uri

    @property
    def DATABRICKS_TOKEN(self) -> str:
        return ""

    def sql(self, statement: str):
        return self.catalog.sql(statement)

    def table(self, name: str):
        return self.catalog.table(name)
