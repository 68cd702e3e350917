#!/bin/bash
# Validation after a kernel change, cheapest diagnostics first.  Every step is bounded by its own timeout and the
# fused-path steps only run if the dgrad numerics case passed, so a real deadlock costs one case timeout, not the call.
#   gpurun --timeout 1500 -- 'bash benchmarks/run_validation.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -u benchmarks/preflight.py > gpurun_out/preflight.log 2>&1; echo "preflight rc=$? (3 = something was stale and got rebuilt)"; cat gpurun_out/preflight.log | grep PREFLIGHT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
CASE_TIMEOUT=${CASE_TIMEOUT:-200} bash benchmarks/run_gpu_checks.sh elementwise conv_dgrad
BT=${BENCH_TIMEOUT:-150}
if grep -q "CASE conv_dgrad PASS" gpurun_out/check_conv_dgrad.log; then
  CASE_TIMEOUT=${CASE_TIMEOUT:-200} bash benchmarks/run_gpu_checks.sh engine
  timeout $BT python -u bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/bench8.log 2>&1; echo "bench8 (default) rc=$?"
  timeout $BT python -u bench.py --steps 20 --warmup 5 --no-e2e --no-fuse-bwd-reduce > gpurun_out/bench8_nofuse.log 2>&1; echo "bench8_nofuse rc=$?"
  timeout $BT python -u bench.py --steps 20 --warmup 5 --no-e2e --overlap-wgrad > gpurun_out/bench8_ov.log 2>&1; echo "bench8_ov rc=$?"
else
  echo "conv_dgrad did NOT pass: skipping every fused-BN-backward step"
  tail -15 gpurun_out/check_conv_dgrad.log
  timeout $BT python -u bench.py --steps 20 --warmup 5 --no-e2e --no-fuse-bwd-reduce > gpurun_out/bench8_nofuse.log 2>&1; echo "bench8_nofuse rc=$?"
fi
for f in gpurun_out/bench8*.log; do echo "== $f"; tail -c 700 $f; echo; done
