#!/bin/bash
# compute-sanitizer passes over the hand-written kernels.  Small shapes only: the tool slows
# kernels 10-100x.  usage (inside gpurun): bash benchmarks/run_sanitizer.sh [memcheck|racecheck|synccheck ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-"memcheck racecheck synccheck"}
for tool in $TOOLS; do
  echo "=== compute-sanitizer --tool $tool" | tee -a gpurun_out/sanitizer.log
  timeout ${TOOL_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python benchmarks/sanitizer_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/sanitizer.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/sanitizer_$tool.log | head -20 | tee -a gpurun_out/sanitizer.log
done
