#!/bin/bash
# Round-2 GPU call 6 (1 GPU): last-CTA BN tails, static prefetch slots (block-gradient timing), GPU JPEG decode, full pytest, A/B tails
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -u benchmarks/preflight.py > $O/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT $O/preflight.log | head -3
for c in tails block_grad fused_infer engine; do
  timeout 400 python -u benchmarks/gpu_check.py $c > $O/check_$c.log 2>&1
  echo "== $c rc=$? $(grep -c PASS $O/check_$c.log) pass / $(grep -E '^CHECK' $O/check_$c.log | grep -c FAIL) fail"
  grep -E "^(CHECK|CASE|TIME|INFO fused)" $O/check_$c.log | grep -E "FAIL|EXCEPTION|TIME|INFO" | head -14
  grep -B2 -A12 "Traceback" $O/check_$c.log | head -30
done
for r in 1 2; do
  timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline > $O/ab6_default_$r.log 2>&1; echo "bench default $r rc=$?"
  B200DDL_TAILS=1 timeout 200 python -u bench.py --steps 40 --warmup 5 --no-e2e --no-baseline > $O/ab6_notails_$r.log 2>&1; echo "bench no-tails $r rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/ab6_*.log')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO JSON', open(f).read()[-600:]); continue
    d = json.loads(ls[-1])
    print(f, f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms launches/step={d['launches_per_step']} sm={d['clocks']['sm_mhz']} {d['clocks']['reasons']} loss={d['loss']:.4f}")
PY
timeout 900 python -u -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 $O/pytest_gpu.log
timeout 200 python -u -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
