#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_1gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_1gpu_tests.log
tail -4 gpurun_out/r2_1gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_1gpu_bench_c4.json 2> gpurun_out/r2_1gpu_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_1gpu_bench_c4.json'))
print('c4', {k:d[k] for k in ('value','ms_per_step','phase_ms')}, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'verify', d['verify'], 'late', d['late_run']['ms_per_step'], d['late_run']['phase_ms'], d['late_run']['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['single_thread']['value'], 'fa2', d['secondary_path_fa2'].get('value'), d['secondary_path_fa2'].get('e2e',{}).get('value'), d['secondary_path_fa2'].get('error'))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reference arm', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
for wl in c3 c2s; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-fa2 > gpurun_out/r2_1gpu_bench_$wl.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_1gpu_bench_$wl.json'))
print('$wl', d['ms_per_step'], d['value'], d['phase_ms'], d['e2e']['value'], d['verify']['ok'], d['late_run']['ms_per_step'])
PY
done
SVI_LS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ring --launch-skip 2 --launch-count 3 -o gpurun_out/r2_1gpu_sweeps_c4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 > gpurun_out/r2_1gpu_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_1gpu_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 > /dev/null 2>&1; echo "ncu list rc=$?"
