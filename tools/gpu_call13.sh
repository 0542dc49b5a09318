#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c13_tests.log
tail -4 gpurun_out/r2c13_tests.log
for wl in c2s c3; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl graph', d['ms_per_step'], d['phase_ms'], d['e2e']['value'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c13_launches_c2s.csv python bench.py --workload c2s --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 > /dev/null 2>&1
# same box A/B at config 4: graph vs plain launches
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 graph', d['ms_per_step'], d['phase_ms'])"
SVI_LS_NO_GRAPH=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 plain', d['ms_per_step'], d['phase_ms'])"
timeout 600 python bench_fa2.py --workload c4 --steps 200 --no-cpu-baseline > gpurun_out/r2c13_fa2_c4.json 2>gpurun_out/r2c13_fa2_c4.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c13_fa2_c4.json'))
print('fa2 c4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['iterations_per_s'])
PY
timeout 600 python bench_fa2.py --workload tiny --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fa2 tiny', d['value'], d['ms_per_step'], 'e2e', d['e2e']['iterations_per_s'])"
