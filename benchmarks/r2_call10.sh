#!/bin/bash
# Round-2 GPU call 10 (1 GPU, last ~3 min): what nvJPEG's GPU-hybrid decode launches (kernel list with durations) and one full
# ncu capture of our batched resize kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/jpeg_launches.csv python benchmarks/jpeg_decode_check.py --batch 256 --iters 1 --backends gpu_hybrid > $O/jpeg_launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/jpeg_launches.csv', errors='replace')) if len(r) > 5]
hdr = next((r for r in rows if 'Kernel Name' in r), None)
if hdr:
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi: continue
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        k = r[ki][:90]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"LAUNCH {n:5d} x {t / n / 1e3:9.1f} us  total {t / 1e6:8.3f} ms  {k}")
PY
timeout 120 ncu --set full --clock-control none --import-source on -k regex:resize_triangle_batched -s 2 -c 1 -f -o $O/ncu_resize_batched python benchmarks/jpeg_decode_check.py --batch 256 --iters 1 --backends gpu_hybrid > $O/ncu_resize.log 2>&1; echo "ncu resize rc=$?"
