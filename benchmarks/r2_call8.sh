#!/bin/bash
# Round-2 GPU call 8 (1 GPU): native nvJPEG decoder (backends, pixels, speed), decode PROCESSES vs threads from real JPEG bytes,
# register-tiled depthwise kernel (numerics + timing), pytest -m gpu of the new tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
timeout 200 python -u benchmarks/jpeg_decode_check.py --batch 256 --stored 500x375 > $O/jpeg_decode_check.log 2>&1; echo "jpeg check rc=$?"
grep -E "JPEG_DECODE|Error" $O/jpeg_decode_check.log | cut -c1-500
timeout 300 python -u benchmarks/gpu_check.py mobilenet > $O/check_mobilenet.log 2>&1
echo "== mobilenet rc=$? $(grep -c PASS $O/check_mobilenet.log) pass / $(grep -E '^CHECK' $O/check_mobilenet.log | grep -c FAIL) fail"
grep -E "^(CHECK|TIME|INFO)" $O/check_mobilenet.log | grep -E "FAIL|EXCEPTION|TIME|INFO" | head -24
grep -B2 -A12 "Traceback" $O/check_mobilenet.log | head -30
timeout 600 python -u -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -12 $O/pytest_gpu.log | cut -c1-300
timeout 500 python -u benchmarks/loader_jpeg_bench.py --images 2048 --stored 500x375 --batch 256 --steps 30 --workers 32 --procs 96 --proc-threads 8 --modes procs,gpu > $O/loader_jpeg2.log 2>&1; echo "loader jpeg rc=$?"
grep -E "dataset:|LOADER_JPEG|Error" $O/loader_jpeg2.log | cut -c1-700
