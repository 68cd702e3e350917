#!/bin/bash
# Round-2 GPU call 3 (1 GPU): stem-backward v2 (scatter), inline-wgrad fix, probes extension, HPO process workers, pyfunc @ batch 256
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
for c in stem_bwd engine_serial_wgrad elementwise umma_probe; do
  timeout 300 python -u benchmarks/gpu_check.py $c > $O/check_$c.log 2>&1
  echo "== $c rc=$? $(grep -c PASS $O/check_$c.log) pass / $(grep -E '^CHECK' $O/check_$c.log | grep -c FAIL) fail"
  grep -E "^(CHECK|CASE|TIME)" $O/check_$c.log | grep -E "FAIL|EXCEPTION|TIME" | head -12
done
for r in 1 2; do
  timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline > $O/ab3_default_$r.log 2>&1; echo "bench default $r rc=$?"
  timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline --stem-bwd-fuse > $O/ab3_stemfuse_$r.log 2>&1; echo "bench stemfuse $r rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/ab3_*.log')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO JSON', open(f).read()[-600:]); continue
    d = json.loads(ls[-1])
    print(f, f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms launches/step={d['launches_per_step']} sm={d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 240 $NCU -k regex:stem_pool_bn_bwd -s 3 -c 1 -f -o $O/ncu_stem_bwd_reduce_v2 python benchmarks/ncu_target.py stem_bwd_reduce s1_1x1_64_256 > $O/ncu_stem_v2.log 2>&1; echo "ncu stem v2 rc=$?"
timeout 900 python -u -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 $O/pytest_gpu.log
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=1024
timeout 300 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 600 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
WORKSHOP_INFER_IMAGES=300000 WORKSHOP_INFER_BATCH=256 timeout 900 python -u examples/part2/03_pyfunc_inference.py > $O/ex_p2_03.log 2>&1; echo "pyfunc example rc=$?"
grep -E "INFERENCE_STATS|scored|Error|error" $O/ex_p2_03.log | cut -c1-600
NUM_EVALS=8 timeout 900 python -u examples/part2/01_hpo_single.py > $O/ex_p2_01.log 2>&1; echo "hpo single rc=$?"; grep -E "best|Error" $O/ex_p2_01.log | head -3 | cut -c1-300
