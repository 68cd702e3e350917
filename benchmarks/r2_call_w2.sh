#!/bin/bash
# Round-2 multi-GPU call (2 GPUs): comm kernels (all-reduce sweep, broadcast, DistributedOptimizer algorithms), 2-GPU bench with
# the same-lease baseline, distributed examples incl. HPO with a persistent rank pool, compute-sanitizer on the comm kernels;
# plus the 1-GPU re-measurements of this iteration (stem backward v2b, block-gradient prefetch, wave-aware tile width).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
W=${WORLD:-2}
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
for c in stem_bwd block_grad engine; do
  timeout 300 python -u benchmarks/gpu_check.py $c > $O/check_$c.log 2>&1
  echo "== $c rc=$? $(grep -c PASS $O/check_$c.log) pass / $(grep -E '^CHECK' $O/check_$c.log | grep -c FAIL) fail"
  grep -E "^(CHECK|CASE|TIME)" $O/check_$c.log | grep -E "FAIL|EXCEPTION|TIME" | head -12
done
timeout 400 python -u benchmarks/gpu_check.py conv_time > $O/check_conv_time.log 2>&1; echo "conv_time rc=$?"; grep TIME $O/check_conv_time.log
for r in 1 2; do
  timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline > $O/ab4_default_$r.log 2>&1; echo "bench default $r rc=$?"
  timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline --stem-bwd-fuse > $O/ab4_stemfuse_$r.log 2>&1; echo "bench stemfuse $r rc=$?"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 benchmarks/allreduce_check.py --medium --max-mb 1024 > $O/allreduce_w$W.log 2>&1; echo "allreduce_check rc=$?"
grep -E "^(f32|bf16) +(1024|65536|1048576|16777216|134217728|1073741824) |broadcast|CTA sweep|DistributedOptimizer|ALLREDUCE|FAIL" $O/allreduce_w$W.log | cut -c1-330
timeout 400 $TR --master-port 29733 benchmarks/fused_update_check.py > $O/fused_update_w$W.log 2>&1; echo "fused_update_check rc=$?"; tail -2 $O/fused_update_w$W.log
timeout 900 $TR --master-port 29641 bench.py --gpus $W --steps 30 --warmup 5 > $O/bench_w$W.log 2>&1; echo "bench w$W rc=$?"
timeout 300 $TR --master-port 29651 bench.py --gpus $W --steps 30 --warmup 5 --no-e2e --no-baseline --algo nvls > $O/bench_w${W}_nvls.log 2>&1; echo "bench w$W nvls rc=$?"
timeout 300 $TR --master-port 29661 bench.py --gpus $W --steps 30 --warmup 5 --no-e2e --no-baseline --algo nccl > $O/bench_w${W}_nccl.log 2>&1; echo "bench w$W nccl rc=$?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/ab4_*.log')) + sorted(glob.glob('gpurun_out/bench_w*.log')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO JSON', open(f).read()[-500:]); continue
    d = json.loads(ls[-1])
    b = d.get('baseline') or {}
    print(f, f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms n={d['n_gpus']} algo={d['config'].get('allreduce')} vs_baseline={d.get('vs_baseline')} "
             f"baseline={b.get('value')} graph={b.get('graph')} e2e={(d.get('e2e') or {}).get('value')} sm={d['clocks']['sm_mhz']} {d['clocks']['reasons']} identical={d.get('params_identical_across_ranks')}")
PY
export B200DDL_HOME=/tmp/ws_gpu WORKSHOP_IMAGES=1024
timeout 300 python -u examples/part1/00_setup.py > $O/ex_p1_00.log 2>&1
timeout 600 python -u examples/part1/01_data_prep.py > $O/ex_p1_01.log 2>&1; echo "data prep rc=$?"
HVD_NP=$W timeout 600 python -u examples/part1/03_train_distributed.py > $O/ex_p1_03_np$W.log 2>&1; echo "example p1/03 np=$W rc=$?"; grep -E "Epoch|img/s" $O/ex_p1_03_np$W.log | tail -3 | cut -c1-200
HVD_NP=$W MAX_EVALS=4 HVD_LOGS=none timeout 900 python -u examples/part2/02_hpo_distributed.py > $O/ex_p2_02_np$W.log 2>&1; echo "hpo distributed np=$W rc=$?"
grep -E "HPO_TIMING" $O/ex_p2_02_np$W.log | cut -c1-1200
HVD_NP=$W MAX_EVALS=2 HVD_LOGS=none HVD_PERSISTENT=0 timeout 900 python -u examples/part2/02_hpo_distributed.py > $O/ex_p2_02_np${W}_fresh.log 2>&1; echo "hpo distributed (fresh ranks) np=$W rc=$?"
grep -E "HPO_TIMING" $O/ex_p2_02_np${W}_fresh.log | cut -c1-700
# compute-sanitizer over the cross-GPU kernels (flag barriers at .sys scope, P2P loads/stores, multimem): memcheck
timeout 420 compute-sanitizer --tool memcheck --target-processes all --log-file $O/sanitizer_comm_w${W}_%p.log \
    $TR --master-port 29671 benchmarks/allreduce_check.py --quick --max-mb 64 > $O/sanitizer_comm_run.log 2>&1; echo "sanitizer memcheck rc=$?"
grep -h "ERROR SUMMARY" $O/sanitizer_comm_w${W}_*.log | sort | uniq -c | head -5
tail -2 $O/sanitizer_comm_run.log | cut -c1-200
