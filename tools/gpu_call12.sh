#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c12_tests.log
tail -4 gpurun_out/r2c12_tests.log
for wl in c4 c3 c2s; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2c12_bench_$wl.json 2> gpurun_out/r2c12_bench_$wl.err; echo "bench $wl rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2c12_bench_$wl.json'))
print('$wl', {k:d[k] for k in ('value','ms_per_step','phase_ms')}, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'verify', d['verify'] and d['verify']['ok'], 'late', d['late_run'] and (d['late_run']['ms_per_step'], d['late_run']['phase_ms'], d['late_run']['roofline']['frac']), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
PY
done
SVI_LS_NO_GRAPH=1 timeout 300 python bench.py --workload c2s --steps 20 --warmup 5 --no-cpu-baseline --no-verify --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2s no graph', d['ms_per_step'])"
for tool in racecheck synccheck; do
  SVI_LS_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool --num-cuda-barriers 512 python tools/sanitize_ring.py > gpurun_out/r2c12_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2c12_$tool.log; tail -3 gpurun_out/r2c12_$tool.log
done
# ncu: launch list of a short bench, then full captures of the two phi launches (lo: no tally, up: tally) of step 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c12_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --converged-frac 0 > /dev/null 2>&1; echo "ncu list rc=$?"
SVI_LS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ring --launch-skip 3 --launch-count 2 -o gpurun_out/r2c12_phi_c4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --converged-frac 0 > gpurun_out/r2c12_ncu.log 2>&1; echo "ncu full rc=$?"
