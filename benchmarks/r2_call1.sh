#!/bin/bash
# Round-2 GPU call 1 (1 GPU): numerics of the new kernels, engine vs torchvision, A/B of the fusions, bench with the
# same-lease baseline, per-kernel step profile.  Every step runs under its own timeout; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $O/smi.txt 2>&1
for c in block_grad head stem_bwd engine engine_unfused_block_grad elementwise conv_dgrad conv_wgrad conv_fwd big_numerics; do
  timeout ${CASE_TIMEOUT:-300} python -u benchmarks/gpu_check.py $c > $O/check_$c.log 2>&1
  echo "== $c rc=$? $(grep -c PASS $O/check_$c.log) pass / $(grep -E '^CHECK' $O/check_$c.log | grep -c FAIL) fail"
  grep -E "^(CHECK|CASE|TIME)" $O/check_$c.log | grep -E "FAIL|EXCEPTION|TIME" | head -12
done
timeout 200 python -u -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
for r in 1 2; do
  for v in default no_block_grad no_stem_bwd_fuse; do
    case $v in default) F="";; no_block_grad) F="--no-block-grad";; no_stem_bwd_fuse) F="--no-stem-bwd-fuse";; esac
    timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline $F > $O/ab_${v}_$r.log 2>&1; echo "bench $v round $r rc=$?"
  done
done
timeout 900 python -u bench.py --steps 30 --warmup 5 > $O/bench_full.log 2>&1; echo "bench full rc=$?"
timeout 300 python -u benchmarks/profile_step.py 256 > $O/profile_step.log 2>&1; echo "profile rc=$?"
python - <<'PY'
import glob, json
def last(f):
    try:
        ls = [l for l in open(f) if l.startswith('{')]
        return json.loads(ls[-1]) if ls else None
    except Exception:
        return None
for f in sorted(glob.glob('gpurun_out/ab_*.log')):
    d = last(f)
    print(f, 'NO JSON' if d is None else f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms launches/step={d['launches_per_step']} sm={d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
d = last('gpurun_out/bench_full.log')
if d is None:
    print('bench_full NO JSON'); print(open('gpurun_out/bench_full.log').read()[-1500:])
else:
    print('bench_full', f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms vs_baseline={d['vs_baseline']} e2e={d['e2e'] and d['e2e']['value']}")
    print(' baseline', json.dumps(d.get('baseline'))[:700])
PY
head -40 $O/step_kernels.txt 2>/dev/null
