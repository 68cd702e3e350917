"""Print which in-tree extensions match their sources, and build the stale ones ONCE up front.

Run this first in every GPU command line: a snapshot taken while an extension was being rebuilt carries new sources with
an old binary, and every later process would then spend its whole timeout recompiling (silently, before this existed).
Exit code 3 if anything had to be rebuilt, so the caller can see it in the log.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from b200ddl.ops import _build  # noqa: E402

stale = [n for n in _build.EXTENSIONS if not _build.is_built(n)]
print("PREFLIGHT extensions:", {n: (n not in stale) for n in _build.EXTENSIONS}, flush=True)
for n in stale:
    t0 = time.time()
    print(f"PREFLIGHT rebuilding stale extension {n} ...", flush=True)
    _build.build(n)
    print(f"PREFLIGHT rebuilt {n} in {time.time() - t0:.0f}s", flush=True)
sys.exit(3 if stale else 0)
