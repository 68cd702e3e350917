#!/bin/bash
# Runs every GPU check in its own process under a timeout; logs land in gpurun_out/.
# usage: bash benchmarks/run_gpu_checks.sh [case ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES=${@:-"elementwise conv_fwd conv_dgrad conv_wgrad conv_time"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
for c in $CASES; do
  echo "=== $c" | tee -a gpurun_out/checks.log
  timeout ${CASE_TIMEOUT:-240} python -u benchmarks/gpu_check.py $c > gpurun_out/check_$c.log 2>&1
  rc=$?
  echo "rc=$rc" >> gpurun_out/check_$c.log
  grep -E "^(CHECK|CASE|TIME)" gpurun_out/check_$c.log | grep -E "FAIL|CASE|TIME" | head -60 | tee -a gpurun_out/checks.log
  echo "exit $rc" | tee -a gpurun_out/checks.log
done
