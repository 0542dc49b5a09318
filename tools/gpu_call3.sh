#!/bin/bash
# round 2, call 3 (1 GPU): sharded path in one process vs the oracle; full gpu suite; bench with checksum
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 600 python -m pytest tests/test_gpu_mg.py -x -q > gpurun_out/r2c3_mg_tests.log 2>&1; echo "mg tests rc=$?" >> gpurun_out/r2c3_mg_tests.log
tail -15 gpurun_out/r2c3_mg_tests.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_mg.py > gpurun_out/r2c3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c3_tests.log
tail -5 gpurun_out/r2c3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --checksum 6 --no-cpu-baseline > gpurun_out/r2c3_bench_c4.json 2> gpurun_out/r2c3_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3_bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','phase_ms','verify','checksum')}, d['roofline']['frac'], d['e2e']['value'])
PY
