#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c14_tests.log
tail -4 gpurun_out/r2c14_tests.log
timeout 300 python tools/probe_phi.py c4 8 2>&1 | tail -1
for wl in c2s c3; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', d['ms_per_step'], d['phase_ms'], d['e2e']['value'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fa2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['phase_ms'], d['roofline']['frac'], d['verify']['ok'], d['late_run']['ms_per_step'], d['late_run']['phase_ms'])"
