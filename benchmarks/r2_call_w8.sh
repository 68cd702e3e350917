#!/bin/bash
# Round-2 8-GPU call (expensive: 8x charge - keep it short): scaling bench, all-reduce sweep, comm-free control,
# config 4 (1 M images through pyfunc.spark_udf), config 3 (HPO over the np=8 distributed trainer, 8 trials).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
W=${WORLD:-8}
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29641 bench.py --gpus $W --steps 30 --warmup 5 --no-baseline > $O/bench_w$W.log 2>&1; echo "bench w$W rc=$?"
timeout 300 $TR --master-port 29651 bench.py --gpus $W --steps 30 --warmup 5 --no-baseline --no-e2e --algo p2p > $O/bench_w${W}_p2p.log 2>&1; echo "bench w$W p2p rc=$?"
# comm-free control: W independent 1-GPU steps on the same box at the same time (what the clocks / power cap alone cost)
for i in $(seq 0 $((W-1))); do
  CUDA_VISIBLE_DEVICES=$i timeout 300 python -u bench.py --steps 30 --warmup 5 --no-e2e --no-baseline > $O/control_w${W}_gpu$i.log 2>&1 &
done
wait; echo "control done"
timeout 400 $TR --master-port 29611 benchmarks/allreduce_check.py --medium --f32-only --max-mb 1024 > $O/allreduce_w$W.log 2>&1; echo "allreduce_check rc=$?"
grep -E "^f32 |broadcast .* MiB|CTA sweep|DistributedOptimizer|ALLREDUCE|FAIL" $O/allreduce_w$W.log | cut -c1-300
python - <<'PY'
import glob, json
def last(f):
    ls = [l for l in open(f) if l.startswith('{')]
    return json.loads(ls[-1]) if ls else None
for f in sorted(glob.glob('gpurun_out/bench_w*.log')) + sorted(glob.glob('gpurun_out/control_w*_gpu*.log')):
    d = last(f)
    if d is None: print(f, 'NO JSON', open(f).read()[-400:]); continue
    print(f, f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms n={d['n_gpus']} algo={d['config'].get('allreduce')} e2e={(d.get('e2e') or {}).get('value')} "
             f"sm={d['clocks']['sm_mhz']} min={d['clocks'].get('sm_mhz_min')} P={d['clocks'].get('power_w_max')} {d['clocks']['reasons']} identical={d.get('params_identical_across_ranks')}")
PY
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=512
timeout 200 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 400 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
WORKSHOP_INFER_IMAGES=${INFER:-1000000} WORKSHOP_INFER_BATCH=256 timeout 600 python -u examples/part2/03_pyfunc_inference.py > $O/ex_p2_03_w$W.log 2>&1; echo "pyfunc example rc=$?"
grep -E "INFERENCE_STATS|Error|error" $O/ex_p2_03_w$W.log | cut -c1-1500
HVD_NP=$W MAX_EVALS=${EVALS:-8} HVD_LOGS=none timeout 600 python -u examples/part2/02_hpo_distributed.py > $O/ex_p2_02_np$W.log 2>&1; echo "hpo distributed np=$W rc=$?"
grep -E "HPO_TIMING|Error" $O/ex_p2_02_np$W.log | cut -c1-2500
