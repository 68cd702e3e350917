"""Rank-process entry point used by `Runner`: unpickle (fn, kwargs), run, hand rank 0's value back.

One-shot mode (``_worker payload result``) serves `Runner.run` with fresh processes.  Serve mode (``_worker --serve``)
keeps the rank alive between jobs - CUDA context, loaded extensions, the NCCL / gloo process group and the symmetric
flag buffers survive from one `run()` to the next, which is what makes back-to-back HPO trials over the distributed
trainer cheap (SURVEY.md hard part 6: "seconds, not minutes").  Protocol on stdin, one line per job:
``<payload path>\t<result path>``; the worker answers ``@@B200DDL_JOB_DONE <code>`` on stdout.  EOF ends the worker.
"""
from __future__ import annotations

import os
import sys
import time
import traceback

DONE = "@@B200DDL_JOB_DONE"


def _run_job(payload: str, result: str, persistent: bool) -> int:
    import cloudpickle

    t0 = time.time()
    with open(payload, "rb") as f:
        fn, kwargs = cloudpickle.load(f)
    rank = int(os.environ.get("RANK", "0"))
    code = 0
    try:
        value = fn(**kwargs)
        if rank == 0:
            with open(result + ".tmp", "wb") as f:
                cloudpickle.dump((True, value, {"worker_job_s": time.time() - t0}), f)
            os.replace(result + ".tmp", result)
    except BaseException:
        tb = traceback.format_exc()
        sys.stdout.write(tb)
        sys.stdout.flush()
        if rank == 0:
            try:
                with open(result + ".tmp", "wb") as f:
                    cloudpickle.dump((False, tb, {}), f)
                os.replace(result + ".tmp", result)
            except Exception:
                pass
        code = 1
    finally:
        try:
            if persistent:
                import gc

                from b200ddl.parallel import symm

                gc.collect()
                symm.end_job()  # keep the process group + flag buffers, drop this job's symmetric data buffers
            else:
                from b200ddl.parallel import core

                core.shutdown()
        except Exception:
            pass
    return code


def main() -> int:
    if len(sys.argv) >= 2 and sys.argv[1] == "--serve":
        # warm start: pay the imports (torch, extensions) once, before the first job arrives
        try:
            import torch  # noqa: F401
            import b200ddl  # noqa: F401
        except Exception:
            pass
        sys.stdout.write(f"{DONE} ready\n")
        sys.stdout.flush()
        for line in sys.stdin:
            line = line.strip()
            if not line:
                continue
            payload, result = line.split("\t")
            code = _run_job(payload, result, persistent=True)
            sys.stdout.write(f"{DONE} {code}\n")
            sys.stdout.flush()
            if code != 0:
                return code  # gang semantics: a failed rank ends; the driver stops the others
        try:
            from b200ddl.parallel import core

            core.shutdown()
        except Exception:
            pass
        return 0
    return _run_job(sys.argv[1], sys.argv[2], persistent=False)


if __name__ == "__main__":
    sys.exit(main())
