#!/bin/bash
# Alternating A/B of two bench configurations on N GPUs (torchrun, 127.0.0.1), reporting throughput AND the cross-rank
# parameter checksum of each run.   gpurun --gpus 2 --timeout 900 -- 'bash benchmarks/run_ab_multi.sh 2 "" "--overlap-wgrad"'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$1; A="$2"; B="$3"; ROUNDS=${ROUNDS:-2}; STEPS=${STEPS:-50}
python -u benchmarks/preflight.py | grep PREFLIGHT
run() {  # tag, port, extra flags
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 \
    bench.py --gpus $N --steps $STEPS --warmup 5 --no-e2e $3 > gpurun_out/abm_$1.log 2>&1
  echo "$1 [$3] rc=$?"
}
port=29610
for r in $(seq 1 $ROUNDS); do
  run A_$r $port "$A"; port=$((port+1))
  run B_$r $port "$B"; port=$((port+1))
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/abm_*.log')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls:
        print(f, 'NO JSON;', 'tail:', open(f).read()[-400:].replace('\n', ' | '))
        continue
    d = json.loads(ls[-1])
    print(f, f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms n_gpus={d['n_gpus']} overlap={d['config']['overlap_wgrad']} "
             f"params_identical={d.get('params_identical_across_ranks')} loss={d['loss']:.4f} sm={d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
