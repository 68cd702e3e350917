#!/usr/bin/env python
"""Top stall locations of an ncu report (source page, SASS view): `python tools/ncu_top_stalls.py file.ncu-rep [N]`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print(f"{rep}: {len(body)} SASS instructions, {tot} samples")
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print("stall reasons: " + ", ".join(f"{k[6:]}={v} ({100 * v / max(tot, 1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
body.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
for r in body[:topn]:
    n = int(r[ci["# Samples"]] or 0)
    why = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{n:7d} {100 * n / max(tot, 1):5.1f}%  {r[ci['Source']].strip()[:70]:70s} exec={r[ci['Instructions Executed']]:>8s}  " +
          " ".join(f"{w}={c}" for c, w in why if c))
