#!/bin/bash
# round 2, call 1: new parity tests, sanitizers on the ring sweeps, bench with --verify, ncu of the s3 sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c1_tests.log
tail -5 gpurun_out/r2c1_tests.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --num-cuda-barriers 64 python tools/sanitize_ring.py > gpurun_out/r2c1_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2c1_$tool.log; tail -3 gpurun_out/r2c1_$tool.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench_c4.json 2> gpurun_out/r2c1_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c1_bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','phase_ms','verify')}, d['roofline']['frac'], d['e2e']['value'])
PY
# ncu --set full of the s3 sweep: launches of k_sweep_ring alternate phi, s3; the 4th is the s3 sweep of step 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ring --launch-skip 3 --launch-count 1 \
  -o gpurun_out/r2c1_s3_c4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r2c1_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/probe_late.py c4 > gpurun_out/r2c1_probe_late.log 2>&1; tail -3 gpurun_out/r2c1_probe_late.log
