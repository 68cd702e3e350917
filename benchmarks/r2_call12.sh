#!/bin/bash
# Round-2 GPU call 12 (1 GPU, the last minute): whole-step A/B of the per-kernel rows-per-iteration choice (3) vs 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for u in 3 1 3 1; do
  B200DDL_BN_UNROLL=$u timeout 40 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline > $O/ab2_bn_unroll_$u.log 2>&1
  echo "bench unroll=$u rc=$? $(grep '^{' $O/ab2_bn_unroll_$u.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read() or '{}'); print(d.get('value'), d.get('ms_per_step'), d.get('clocks',{}).get('sm_mhz'), d.get('config',{}).get('bn_rows_unroll'), d.get('loss'))" 2>&1 | tail -1)"
  cp $O/ab2_bn_unroll_$u.log $O/ab2_bn_unroll_${u}_$RANDOM.log
done
