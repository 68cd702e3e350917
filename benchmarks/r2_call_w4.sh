#!/bin/bash
# Round-2 4-GPU call: the missing 4-GPU all-reduce sweep (BASELINE.json config 5), 4-GPU bench, gated multi-GPU pytest
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 benchmarks/allreduce_check.py --medium --f32-only --max-mb 1024 > $O/allreduce_w4.log 2>&1; echo "allreduce_check rc=$?"
grep -E "^f32 |broadcast .* MiB|CTA sweep|DistributedOptimizer|ALLREDUCE|FAIL" $O/allreduce_w4.log | cut -c1-300
timeout 700 $TR --master-port 29641 bench.py --gpus 4 --steps 30 --warmup 5 --baseline-timeout 150 > $O/bench_w4.log 2>&1; echo "bench w4 rc=$?"
grep '^{' $O/bench_w4.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); b = d.get('baseline') or {}; print('bench4', round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('allreduce'), 'e2e', (d.get('e2e') or {}).get('value'), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('params_identical_across_ranks'), 'vs_baseline', d.get('vs_baseline'), 'baseline', b.get('value'), b.get('graph'), str(b.get('unavailable'))[:300], str(b.get('notes'))[:300])"
timeout 600 python -u -m pytest tests/test_gpu_kernels.py -q -m gpu -k "multi_gpu or two_gpus" > $O/pytest_multi_gpu_w4.log 2>&1; echo "pytest multi-gpu rc=$?"; tail -3 $O/pytest_multi_gpu_w4.log
