#!/bin/bash
# Round-2 GPU call 9 (1 GPU, final tree): where CPU decode stops scaling on this host, training from real JPEG bytes with the
# reworked decode-process loader and the native nvJPEG path, the headline bench with the same-lease baseline, pytest -m gpu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
timeout 300 python -u benchmarks/cpu_decode_scaling.py > $O/cpu_decode_scaling.log 2>&1; echo "cpu scaling rc=$?"
grep -E "CPU_DECODE|Error" $O/cpu_decode_scaling.log | cut -c1-400
timeout 400 python -u benchmarks/loader_jpeg_bench.py --images 1024 --stored 500x375 --batch 256 --steps 30 --procs 96 --proc-threads 4 --modes procs,gpu > $O/loader_jpeg3.log 2>&1; echo "loader jpeg rc=$?"
grep -E "dataset:|LOADER_JPEG|Error" $O/loader_jpeg3.log | cut -c1-700
timeout 500 python -u bench.py > $O/bench_final2.log 2>&1; echo "bench rc=$?"; grep '^{' $O/bench_final2.log | tail -1 | cut -c1-900
timeout 600 python -u -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-300
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:dwconv3x3_tiled -s 3 -c 1 -f -o $O/ncu_dwconv_tiled_56_256 python benchmarks/ncu_target.py dwconv s1_1x1_256_64 > $O/ncu_dwconv_tiled.log 2>&1; echo "ncu dwconv tiled rc=$?"
