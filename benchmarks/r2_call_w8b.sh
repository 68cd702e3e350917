#!/bin/bash
# Round-2 second 8-GPU call: ONLY config 4 (1 M images through pyfunc.spark_udf) after the worker warm-up / Arrow-result change.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
W=${WORLD:-8}
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=512
timeout 200 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 300 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
WORKSHOP_INFER_IMAGES=${INFER:-1000000} WORKSHOP_INFER_FRAG_ROWS=${FRAG_ROWS:-4096,4096,8192} WORKSHOP_INFER_BATCH=256 timeout 400 python -u examples/part2/03_pyfunc_inference.py > $O/ex_p2_03_w${W}b.log 2>&1; echo "pyfunc example rc=$?"
grep -E "INFERENCE_STATS|Error|error" $O/ex_p2_03_w${W}b.log | cut -c1-2500
