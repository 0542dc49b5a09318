#!/bin/bash
# round 2, call 2: partitioned neighbour lists + one arg-max per link + device CSR build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
tail -5 gpurun_out/r2c2_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench_c4.json 2> gpurun_out/r2c2_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c2_bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','phase_ms','verify','setup_s')}, d['roofline']['frac'], d['e2e']['value'])
PY
timeout 300 python tools/probe_phi.py c4 8 > gpurun_out/r2c2_probe_phi.log 2>&1; tail -2 gpurun_out/r2c2_probe_phi.log
timeout 300 python tools/probe_late.py c4 > gpurun_out/r2c2_probe_late.log 2>&1; tail -3 gpurun_out/r2c2_probe_late.log
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_c3.json 2> gpurun_out/r2c2_bench_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c2_bench_c3.json'))
print('c3', {k:d[k] for k in ('value','ms_per_step','phase_ms','verify')})
PY
