#!/bin/bash
# K > 1024 / FastAMM2 K > 512 on a B200: the device tests of the block-per-row kernels (first full hardware run of
# tests/test_gpu_wide.py's later cases and of tests/test_gpu_wide_fa2.py), then a first timing of the wide
# link-sampling path (bench.py --workload widek: n=50 000, K=2048, 1e6 links).  gpurun --timeout 900 -- tools/gpu_wide.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_wide_fa2.py -v -m gpu > gpurun_out/wide_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/wide_tests.log
grep -E "PASSED|FAILED|ERROR|passed|failed|rc=" gpurun_out/wide_tests.log | tail -40
timeout 600 python bench.py --workload widek --steps 10 --warmup 3 --no-cpu-baseline --no-fa2 > gpurun_out/wide_bench.json 2> gpurun_out/wide_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/wide_bench.json'))
print('widek', d['ms_per_step'], d['value'], d['phase_ms'], 'frac', d['roofline']['frac'], 'verify', d['verify'])
PY
