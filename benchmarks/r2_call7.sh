#!/bin/bash
# Round-2 GPU call 7 (1 GPU, final tree): pytest -m gpu, smoke, bench with the same-lease baseline, reference arm, training from REAL
# JPEG bytes (cpu / nvJPEG decode), ncu of the two kernel families added late (inference epilogue, depthwise 3x3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
timeout 600 python -u -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 $O/pytest_gpu.log
timeout 200 python -u -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 500 python -u bench.py > $O/bench_final.log 2>&1; echo "bench rc=$?"; grep '^{' $O/bench_final.log | tail -1 | cut -c1-1200
timeout 60 python -u bench.py --impl reference > $O/bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -1 $O/bench_reference.log | cut -c1-300
timeout 400 python -u benchmarks/loader_jpeg_bench.py --images 2048 --stored 500x375 --batch 256 --steps 30 --workers 32 > $O/loader_jpeg.log 2>&1; echo "loader jpeg rc=$?"
grep -E "dataset:|LOADER_JPEG|Error" $O/loader_jpeg.log | cut -c1-700
NCU="ncu --set full --clock-control none --import-source on"
timeout 240 $NCU -k regex:conv_igemm -s 3 -c 1 -f -o $O/ncu_fwd_infer_s1_1x1_64_256 python benchmarks/ncu_target.py fwd_infer s1_1x1_64_256 > $O/ncu_fwd_infer.log 2>&1; echo "ncu fwd_infer rc=$?"
timeout 240 $NCU -k regex:dwconv3x3 -s 3 -c 1 -f -o $O/ncu_dwconv_56_256 python benchmarks/ncu_target.py dwconv s1_1x1_256_64 > $O/ncu_dwconv.log 2>&1; echo "ncu dwconv rc=$?"
ls -la $O/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
