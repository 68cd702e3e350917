"""Worker processes started as `python -m <module>` and driven over an authenticated unix-socket connection.

Why not `multiprocessing.Process` (spawn)?  Spawned children re-import the parent's `__main__` - and the workshop scripts
are notebook-style: no `if __name__ == "__main__"` guard - so a spawn from inside an example would re-run the whole
script in every child (Python refuses: "An attempt has been made to start a new process before the current process has
finished its bootstrapping phase").  A plain subprocess runs the framework's own worker module instead; objects travel as
pickles over the connection (cloudpickle by value for user functions / classes defined in `__main__`)."""
from __future__ import annotations

import os
import secrets
import subprocess
import sys
import tempfile
from multiprocessing.connection import Client, Connection, Listener
from typing import List, Optional, Tuple


def start_worker(module: str, slot: int, env: Optional[dict] = None, timeout_s: float = 180.0) -> Tuple[subprocess.Popen, Connection]:
    """Start `python -m module <address> <slot>`; returns (process, connection).  The child calls `connect_parent()`."""
    key = secrets.token_bytes(16)
    addr = os.path.join(tempfile.mkdtemp(prefix="b200ddl_pp_"), "sock")
    listener = Listener(addr, family="AF_UNIX", authkey=key)
    e = dict(os.environ)
    e.update(env or {})
    e["B200DDL_PP_KEY"] = key.hex()
    # the child must be able to import whatever the parent can (user modules that pickles refer to by name)
    e["PYTHONPATH"] = os.pathsep.join([p for p in sys.path if p] + [e.get("PYTHONPATH", "")])
    proc = subprocess.Popen([sys.executable, "-m", module, addr, str(slot)], env=e, stdin=subprocess.DEVNULL,
                            cwd=os.getcwd(), start_new_session=True)
    listener._listener._socket.settimeout(timeout_s)
    try:
        conn = listener.accept()
    except Exception:
        proc.kill()
        raise RuntimeError(f"worker {module}[{slot}] did not connect within {timeout_s}s (exit code {proc.poll()})") from None
    finally:
        listener.close()
    return proc, conn


def connect_parent() -> Tuple[Connection, int]:
    """In the child: (connection to the parent, slot)."""
    addr, slot = sys.argv[1], int(sys.argv[2])
    return Client(addr, family="AF_UNIX", authkey=bytes.fromhex(os.environ["B200DDL_PP_KEY"])), slot


def start_script_workers(script: str, n: int, env: Optional[dict] = None, timeout_s: float = 120.0) -> List[Tuple[subprocess.Popen, Connection]]:
    """Start `n` copies of `python <script> <address> <slot>` AT ONCE (all processes are launched before the first
    connection is awaited: ~100 light workers come up in about the time of one).  The script is run by PATH, not as a
    module of the package, so a worker that only needs numpy / PIL does not pay the framework's (torch) import."""
    e = dict(os.environ)
    e.update(env or {})
    started = []
    for slot in range(n):
        key = secrets.token_bytes(16)
        addr = os.path.join(tempfile.mkdtemp(prefix="b200ddl_pp_"), "sock")
        listener = Listener(addr, family="AF_UNIX", authkey=key)
        ee = dict(e)
        ee["B200DDL_PP_KEY"] = key.hex()
        proc = subprocess.Popen([sys.executable, script, addr, str(slot)], env=ee, stdin=subprocess.DEVNULL,
                                start_new_session=True)
        started.append((proc, listener))
    out: List[Tuple[subprocess.Popen, Connection]] = []
    try:
        for proc, listener in started:
            listener._listener._socket.settimeout(timeout_s)
            out.append((proc, listener.accept()))
    except Exception:
        for proc, _ in started:
            proc.kill()
        raise RuntimeError(f"a worker of {os.path.basename(script)} did not connect within {timeout_s}s") from None
    finally:
        for _, listener in started:
            listener.close()
    return out
