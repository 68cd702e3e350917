"""Rank-process entry point used by `Runner`: unpickle (fn, kwargs), run, hand rank 0's value back."""
from __future__ import annotations

import os
import sys
import traceback


def main() -> int:
    payload, result = sys.argv[1], sys.argv[2]
    import cloudpickle

    with open(payload, "rb") as f:
        fn, kwargs = cloudpickle.load(f)
    rank = int(os.environ.get("RANK", "0"))
    code = 0
    try:
        value = fn(**kwargs)
        if rank == 0:
            with open(result + ".tmp", "wb") as f:
                cloudpickle.dump((True, value), f)
            os.replace(result + ".tmp", result)
    except BaseException:
        tb = traceback.format_exc()
        sys.stdout.write(tb)
        sys.stdout.flush()
        if rank == 0:
            try:
                with open(result + ".tmp", "wb") as f:
                    cloudpickle.dump((False, tb), f)
                os.replace(result + ".tmp", result)
            except Exception:
                pass
        code = 1
    finally:
        try:
            from b200ddl.parallel import core

            core.shutdown()
        except Exception:
            pass
    return code


if __name__ == "__main__":
    sys.exit(main())
