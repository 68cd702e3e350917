#!/bin/bash
# Round-2 GPU call 2 (1 GPU): full pytest -m gpu, bench with the same-lease baseline, ncu captures of the hot kernels,
# clean (no side-stream overlap) per-kernel profile, pyfunc inference through the API, HPO examples.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
timeout 900 python -u -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python -u bench.py --steps 30 --warmup 5 > $O/bench_full.log 2>&1; echo "bench full rc=$?"
grep '^{' $O/bench_full.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms vs_baseline', d['vs_baseline'], 'e2e', d['e2e'] and round(d['e2e']['value'],1))
    print(' baseline', json.dumps(d.get('baseline'))[:900])
"
tail -3 $O/bench_full.log | cut -c1-400
# ---- ncu: one capture per kernel family (full set, source correlation)
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name kernel-regex skip what shape
  timeout 240 $NCU -k regex:$2 -s $3 -c 1 -f -o $O/ncu_$1 python benchmarks/ncu_target.py $4 $5 > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap block_grad_s1 conv_igemm_kernel 3 block_grad s1_1x1_256_64
cap stem_bwd_reduce stem_pool_bn_bwd 3 stem_bwd_reduce s1_1x1_64_256
cap stem_bwd_apply stem_pool_bn_bwd 3 stem_bwd_apply s1_1x1_64_256
cap bn_bwd_apply_s1 bn_bwd_apply 3 bn_bwd_apply s1_1x1_256_64
cap fwd_stats_s2_3x3_128 conv_igemm_kernel 3 fwd_stats s2_3x3_128
cap fwd_stats_s4_3x3_512 conv_igemm_kernel 3 fwd_stats s4_3x3_512
cap wgrad_s3_3x3_256 conv_wgrad_kernel 3 wgrad s3_3x3_256
cap wgrad_s1_1x1_64_256 conv_wgrad_kernel 3 wgrad s1_1x1_64_256
ls -la $O/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
# ---- clean per-kernel profile (wgrad in line, so durations are not inflated by co-running kernels)
timeout 300 python -u benchmarks/profile_step.py 256 inline > $O/profile_step_inline.log 2>&1; echo "profile inline rc=$?"
cp $O/step_kernels.txt $O/step_kernels_inline.txt 2>/dev/null
# ---- config 4 through the API on one GPU: train + package + spark_udf over a lazily generated 200k-image table
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=1024
timeout 300 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 600 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
WORKSHOP_INFER_IMAGES=200000 timeout 900 python -u examples/part2/03_pyfunc_inference.py > $O/ex_p2_03.log 2>&1; echo "pyfunc example rc=$?"
grep -E "INFERENCE_STATS|scored|Error|error" $O/ex_p2_03.log | cut -c1-700
# ---- HPO: single-node trials in 4 worker processes (SparkTrials(parallelism=4)), then distributed trials (np=1 here)
NUM_EVALS=8 timeout 900 python -u examples/part2/01_hpo_single.py > $O/ex_p2_01.log 2>&1; echo "hpo single rc=$?"; tail -3 $O/ex_p2_01.log | cut -c1-300
HVD_NP=1 MAX_EVALS=3 HVD_LOGS=none timeout 900 python -u examples/part2/02_hpo_distributed.py > $O/ex_p2_02.log 2>&1; echo "hpo distributed rc=$?"
grep -E "HPO_TIMING" $O/ex_p2_02.log | cut -c1-900
head -45 $O/step_kernels_inline.txt 2>/dev/null
