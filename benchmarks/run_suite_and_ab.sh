#!/bin/bash
# Full GPU test suite + smoke + an ALTERNATING A/B of two bench configurations (single samples of a 20 ms step differ by
# ~1 % from clock dips alone, so a default is only changed on an interleaved repeat).
#   gpurun --timeout 1500 -- 'bash benchmarks/run_suite_and_ab.sh "" "--overlap-wgrad"'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
A="$1"; B="$2"; ROUNDS=${ROUNDS:-3}; STEPS=${STEPS:-50}
python -u benchmarks/preflight.py > gpurun_out/preflight.log 2>&1; echo "preflight rc=$?"; grep PREFLIGHT gpurun_out/preflight.log
timeout 700 python -u -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 200 python -u -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for r in $(seq 1 $ROUNDS); do
  timeout 150 python -u bench.py --steps $STEPS --warmup 5 --no-e2e $A > gpurun_out/ab_A_$r.log 2>&1; echo "A[$A] round $r rc=$?"
  timeout 150 python -u bench.py --steps $STEPS --warmup 5 --no-e2e $B > gpurun_out/ab_B_$r.log 2>&1; echo "B[$B] round $r rc=$?"
done
timeout 300 python -u bench.py --steps $STEPS --warmup 5 > gpurun_out/bench_default_e2e.log 2>&1; echo "default+e2e rc=$?"
python - <<'PY'
import glob, json
def last(f):
    ls = [l for l in open(f) if l.startswith('{')]
    return json.loads(ls[-1]) if ls else None
for tag in ('A', 'B'):
    for f in sorted(glob.glob(f'gpurun_out/ab_{tag}_*.log')):
        d = last(f)
        print(tag, f, 'NO JSON' if d is None else f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms sm={d['clocks']['sm_mhz']} min={d['clocks'].get('sm_mhz_min')} {d['clocks']['reasons']}")
d = last('gpurun_out/bench_default_e2e.log')
print('default+e2e', 'NO JSON' if d is None else f"{d['value']:.1f} img/s {d['ms_per_step']:.3f} ms  e2e={d['e2e']}")
PY
