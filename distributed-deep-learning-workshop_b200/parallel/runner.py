"""Runner: the HorovodRunner replacement (reference P1/03:391-417, P2/02:302-307; mechanism P1/03:259-263).

`Runner(np).run(fn, **kwargs)` ships `fn` (cloudpickle, *by value*, so closures over driver globals such as
BATCH_SIZE / converters / run ids travel with it - SURVEY.md §5.6) to `np` freshly spawned rank processes on this
node, one per GPU, wires the torchrun-style rendezvous environment (127.0.0.1 TCPStore), streams the workers' logs
back (``driver_log_verbosity='all'``) and returns rank 0's return value.

* ``np=-1`` (or 1): one local process, `size()==1` - the "test on the driver" rung.
* Gang semantics (Spark barrier mode in the reference): if any rank dies, all are killed and the failing rank's
  traceback is raised on the driver.
"""
from __future__ import annotations

import os
import socket
import subprocess
import sys
import tempfile
import threading
import time
from typing import Any, Callable, List, Optional


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class RunnerError(RuntimeError):
    pass


class Runner:
    def __init__(self, np: int = -1, driver_log_verbosity: str = "all", timeout_s: Optional[float] = None,
                 force_cpu: bool = False, env: Optional[dict] = None):
        self.np = 1 if np in (-1, 0, 1) else int(np)
        if self.np < 1:
            raise ValueError("np must be -1 or a positive integer")
        self.driver_log_verbosity = driver_log_verbosity
        self.timeout_s = timeout_s
        self.force_cpu = force_cpu
        self.extra_env = dict(env or {})
        self.last_logs: List[List[str]] = []

    def run(self, main: Callable[..., Any], **kwargs) -> Any:
        import cloudpickle

        repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        with tempfile.TemporaryDirectory(prefix="b200ddl_run_") as tmp:
            payload = os.path.join(tmp, "payload.pkl")
            result = os.path.join(tmp, "result.pkl")
            with open(payload, "wb") as f:
                cloudpickle.dump((main, kwargs), f)
            port = _free_port()
            procs: List[subprocess.Popen] = []
            logs: List[List[str]] = [[] for _ in range(self.np)]
            threads = []
            for r in range(self.np):
                env = dict(os.environ)
                env.update(self.extra_env)
                env.update({
                    "RANK": str(r), "LOCAL_RANK": str(r), "WORLD_SIZE": str(self.np),
                    "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port),
                    "B200DDL_RUNNER": "1", "PYTHONUNBUFFERED": "1",
                    "PYTHONPATH": repo_root + os.pathsep + env.get("PYTHONPATH", ""),
                })
                if self.force_cpu:
                    env["B200DDL_FORCE_CPU"] = "1"
                    env["CUDA_VISIBLE_DEVICES"] = ""
                # N ranks with full-size OpenMP teams oversubscribe the host (measured: 50x slower CPU steps once the
                # spinning teams exceed the cores); like torchrun, give each rank its share unless the user chose.
                if not env.get("OMP_NUM_THREADS"):
                    env["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // self.np))
                p = subprocess.Popen([sys.executable, "-m", "b200ddl.parallel._worker", payload, result],
                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, text=True,
                                     cwd=os.getcwd(), start_new_session=True)
                procs.append(p)
                t = threading.Thread(target=self._pump, args=(p, r, logs[r]), daemon=True)
                t.start()
                threads.append(t)
            failed = None
            t0 = time.time()
            try:
                while True:
                    alive = False
                    for r, p in enumerate(procs):
                        rc = p.poll()
                        if rc is None:
                            alive = True
                        elif rc != 0 and failed is None:
                            failed = (r, rc)
                    if failed is not None or not alive:
                        break
                    if self.timeout_s is not None and time.time() - t0 > self.timeout_s:
                        failed = (-1, -1)
                        break
                    time.sleep(0.05)
            finally:
                if failed is not None:
                    for p in procs:  # gang failure: stop exactly the processes we started
                        if p.poll() is None:
                            p.kill()
                for p in procs:
                    try:
                        p.wait(timeout=30)
                    except Exception:
                        p.kill()
                for t in threads:
                    t.join(timeout=5)
            self.last_logs = logs
            if failed is not None:
                r, rc = failed
                if r < 0:
                    raise RunnerError(f"Runner timed out after {self.timeout_s}s; all ranks were stopped")
                tail = "".join(logs[r][-60:])
                raise RunnerError(f"rank {r} exited with code {rc}; all ranks were stopped.\n--- rank {r} log ---\n{tail}")
            if not os.path.exists(result):
                raise RunnerError("rank 0 finished without producing a result")
            with open(result, "rb") as f:
                ok, value = cloudpickle.load(f)
            if not ok:
                raise RunnerError(f"rank 0 raised:\n{value}")
            return value

    def _pump(self, p: subprocess.Popen, r: int, sink: List[str]) -> None:
        for line in p.stdout:
            sink.append(line)
            if self.driver_log_verbosity == "all":
                sys.stdout.write(f"[rank {r}] {line}")
                sys.stdout.flush()


HorovodRunner = Runner  # drop-in name for notebook parity
