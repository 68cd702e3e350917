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

from ..utils.cpus import usable_cpus


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class RunnerError(RuntimeError):
    pass


class Runner:
    """``persistent=True`` keeps the rank processes (CUDA context, loaded extensions, process group, symmetric flag
    buffers) alive between `run()` calls - the re-initialisation cost of an HPO trial over the distributed trainer drops
    from process spawn + imports + NCCL bootstrap to shipping a pickle.  Use as a context manager or call `close()`.
    `last_timing` holds the driver-side breakdown of the most recent `run()`."""

    def __init__(self, np: int = -1, driver_log_verbosity: str = "all", timeout_s: Optional[float] = None,
                 force_cpu: bool = False, env: Optional[dict] = None, persistent: bool = False):
        self.np = 1 if np in (-1, 0, 1) else int(np)
        if self.np < 1:
            raise ValueError("np must be -1 or a positive integer")
        self.driver_log_verbosity = driver_log_verbosity
        self.timeout_s = timeout_s
        self.force_cpu = force_cpu
        self.extra_env = dict(env or {})
        self.last_logs: List[List[str]] = []
        self.persistent = persistent
        self.last_timing: dict = {}
        self._pool: Optional["_RankPool"] = None

    def __enter__(self) -> "Runner":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def close(self) -> None:
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rank_env(self, r: int, port: int) -> dict:
        repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
            env["OMP_NUM_THREADS"] = str(max(1, usable_cpus() // self.np))   # the cgroup quota, not cpu_count
        return env

    def run(self, main: Callable[..., Any], **kwargs) -> Any:
        if self.persistent:
            return self._run_persistent(main, kwargs)
        return self._run_fresh(main, kwargs)

    def _run_persistent(self, main, kwargs) -> Any:
        import cloudpickle

        t0 = time.time()
        started = False
        if self._pool is None or not self._pool.alive():
            if self._pool is not None:
                self._pool.close()
            self._pool = _RankPool(self)
            started = True
        t1 = time.time()
        with tempfile.TemporaryDirectory(prefix="b200ddl_run_") as tmp:
            payload = os.path.join(tmp, "payload.pkl")
            result = os.path.join(tmp, "result.pkl")
            with open(payload, "wb") as f:
                cloudpickle.dump((main, kwargs), f)
            try:
                failed = self._pool.execute(payload, result, self.timeout_s)
            finally:
                self.last_logs = self._pool.take_logs()
            if failed is not None:
                self._pool.close(kill=True)
                self._pool = None
                r, rc = failed
                if r < 0:
                    raise RunnerError(f"Runner timed out after {self.timeout_s}s; all ranks were stopped")
                tail = "".join(self.last_logs[r][-60:])
                raise RunnerError(f"rank {r} exited with code {rc}; all ranks were stopped.\n--- rank {r} log ---\n{tail}")
            if not os.path.exists(result):
                raise RunnerError("rank 0 finished without producing a result")
            with open(result, "rb") as f:
                rec = cloudpickle.load(f)
        ok, value = rec[0], rec[1]
        self.last_timing = {"pool_started": started, "pool_start_s": t1 - t0, "job_s": time.time() - t1,
                            **(rec[2] if len(rec) > 2 else {})}
        if not ok:
            raise RunnerError(f"rank 0 raised:\n{value}")
        return value

    def _run_fresh(self, main: Callable[..., Any], kwargs) -> Any:
        import cloudpickle

        t_begin = time.time()
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
                env = self._rank_env(r, port)
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
                rec = cloudpickle.load(f)
            ok, value = rec[0], rec[1]
            total = time.time() - t_begin
            wj = (rec[2] if len(rec) > 2 else {}).get("worker_job_s")
            self.last_timing = {"pool_started": True, "total_s": total, "worker_job_s": wj,
                                "spawn_and_import_s": (total - wj) if wj is not None else None}
            if not ok:
                raise RunnerError(f"rank 0 raised:\n{value}")
            return value

    def _pump(self, p: subprocess.Popen, r: int, sink: List[str]) -> None:
        for line in p.stdout:
            sink.append(line)
            if self.driver_log_verbosity == "all":
                sys.stdout.write(f"[rank {r}] {line}")
                sys.stdout.flush()


class _RankPool:
    """`np` long-lived rank processes in serve mode (see parallel/_worker.py)."""

    def __init__(self, runner: Runner):
        from ._worker import DONE

        self.DONE = DONE
        self.runner = runner
        self.np = runner.np
        port = _free_port()
        self.procs: List[subprocess.Popen] = []
        self.logs: List[List[str]] = [[] for _ in range(self.np)]
        self.done_codes: List[List[str]] = [[] for _ in range(self.np)]
        self.cv = threading.Condition()
        for r in range(self.np):
            p = subprocess.Popen([sys.executable, "-m", "b200ddl.parallel._worker", "--serve"], stdin=subprocess.PIPE,
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=runner._rank_env(r, port),
                                 text=True, cwd=os.getcwd(), start_new_session=True)
            self.procs.append(p)
            threading.Thread(target=self._pump, args=(p, r), daemon=True).start()
        self._wait_all(300.0)  # the "ready" line of every rank

    def _pump(self, p: subprocess.Popen, r: int) -> None:
        for line in p.stdout:
            if line.startswith(self.DONE):
                with self.cv:
                    self.done_codes[r].append(line.split()[1])
                    self.cv.notify_all()
                continue
            self.logs[r].append(line)
            if self.runner.driver_log_verbosity == "all":
                sys.stdout.write(f"[rank {r}] {line}")
                sys.stdout.flush()
        with self.cv:
            self.cv.notify_all()

    def alive(self) -> bool:
        return bool(self.procs) and all(p.poll() is None for p in self.procs)

    def _wait_all(self, timeout_s: Optional[float]):
        """Wait until every rank has reported a job end; returns None or (rank, code) of the first failure / (-1, -1)."""
        t0 = time.time()
        with self.cv:
            while True:
                for r, p in enumerate(self.procs):
                    if self.done_codes[r] and self.done_codes[r][0] not in ("0", "ready"):
                        return (r, int(self.done_codes[r][0]))
                    if p.poll() is not None and not self.done_codes[r]:
                        return (r, p.returncode)
                if all(self.done_codes[r] for r in range(self.np)):
                    for r in range(self.np):
                        self.done_codes[r].pop(0)
                    return None
                if timeout_s is not None and time.time() - t0 > timeout_s:
                    return (-1, -1)
                self.cv.wait(timeout=0.1)

    def execute(self, payload: str, result: str, timeout_s: Optional[float]):
        for p in self.procs:
            p.stdin.write(f"{payload}\t{result}\n")
            p.stdin.flush()
        return self._wait_all(timeout_s)

    def take_logs(self) -> List[List[str]]:
        out = self.logs
        self.logs = [[] for _ in range(self.np)]
        return out

    def close(self, kill: bool = False) -> None:
        for p in self.procs:
            try:
                if kill and p.poll() is None:
                    p.kill()  # gang failure: stop exactly the processes we started
                elif p.poll() is None and p.stdin:
                    p.stdin.close()  # EOF: the worker shuts its process group down and exits
            except Exception:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=15)
            except Exception:
                p.kill()
        self.procs = []


HorovodRunner = Runner  # drop-in name for notebook parity
