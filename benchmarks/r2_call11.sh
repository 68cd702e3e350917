#!/bin/bash
# Round-2 GPU call 11 (1 GPU, the last ~2 minutes): BatchNorm apply kernels with 2 / 4 rows per loop iteration - bit-identity
# against the 1-row kernels, per-shape timings, and the whole step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python -u benchmarks/bn_unroll_check.py > $O/bn_unroll_check.log 2>&1; echo "bn unroll check rc=$?"
grep BN_UNROLL $O/bn_unroll_check.log | grep -E "all_identical|\"M\": 802816|\"M\": 50176, \"C\": 1024" | cut -c1-330
for u in 2 1 4; do
  B200DDL_BN_UNROLL=$u timeout 60 python -u bench.py --steps 30 --warmup 5 --no-e2e --no-baseline > $O/ab_bn_unroll_$u.log 2>&1
  echo "bench unroll=$u rc=$? $(grep '^{' $O/ab_bn_unroll_$u.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read() or '{}'); print(d.get('value'), d.get('ms_per_step'), d.get('clocks',{}).get('sm_mhz'), d.get('config',{}).get('bn_rows_unroll'), d.get('loss'))" 2>&1 | tail -1)"
done
