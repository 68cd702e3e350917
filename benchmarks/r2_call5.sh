#!/bin/bash
# Round-2 GPU call 5 (1 GPU): inference epilogue + fused-inference engine, MobileNetV2 engine, GPU JPEG decode, reverted
# block-gradient prefetch, full pytest, bench with baseline child, pyfunc inference through the API
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
for c in fused_infer mobilenet block_grad; do
  timeout 400 python -u benchmarks/gpu_check.py $c > $O/check_$c.log 2>&1
  echo "== $c rc=$? $(grep -c PASS $O/check_$c.log) pass / $(grep -E '^CHECK' $O/check_$c.log | grep -c FAIL) fail"
  grep -E "^(CHECK|CASE|TIME|INFO)" $O/check_$c.log | grep -E "FAIL|EXCEPTION|TIME|INFO" | head -14
  grep -B2 -A12 "Traceback" $O/check_$c.log | head -40
done
timeout 900 python -u -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 $O/pytest_gpu.log
timeout 600 python -u bench.py --steps 30 --warmup 5 > $O/bench_full.log 2>&1; echo "bench full rc=$?"
grep '^{' $O/bench_full.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); b = d.get('baseline') or {}
    print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms vs_baseline', d['vs_baseline'], 'baseline', b.get('value'), b.get('graph'), b.get('unavailable'), 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=512
timeout 300 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 600 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
WORKSHOP_INFER_IMAGES=400000 WORKSHOP_INFER_BATCH=256 timeout 900 python -u examples/part2/03_pyfunc_inference.py > $O/ex_p2_03.log 2>&1; echo "pyfunc example rc=$?"
grep -E "INFERENCE_STATS|scored|Error|error" $O/ex_p2_03.log | cut -c1-600
