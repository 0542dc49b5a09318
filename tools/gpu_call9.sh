#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1200 python -m pytest tests/test_gpu_mg.py tests/test_gpu_parity_ring.py tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -k "not fa2" > gpurun_out/r2c9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c9_tests.log
tail -6 gpurun_out/r2c9_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench_c4.json 2> gpurun_out/r2c9_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','phase_ms')}, d['roofline']['frac'], d['e2e']['value'], d['verify']['ok'])
PY
timeout 300 python tools/probe_phi.py c4 8 2>&1 | tail -1
