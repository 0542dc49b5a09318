#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --checksum 6 > gpurun_out/r2c11_bench_n8.json 2> gpurun_out/r2c11_bench_n8.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c11_bench_n8.json').read().strip().split('\n')[-1])
    print({k:d[k] for k in ('value','ms_per_step','mg_phase_ms','checksum','setup_s')}, 'e2e', d['e2e']['value'], 'verify', d['verify'] and d['verify']['ok'])
except Exception as e: print("parse failed", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_mg.py c4 \
   push=ce_multi,chunks=4,ratio=0.6 push=ce_multi,chunks=3,ratio=0.5 push=ce_multi,chunks=5,ratio=0.6 push=ce_multi,chunks=4,ratio=0.5 push=ce_multi,chunks=6,ratio=0.6 push=sm,chunks=4,ratio=0.6,blocks=96 push=ce,chunks=4,ratio=0.6 \
   > gpurun_out/r2c11_probe_mg_n$N.log 2> gpurun_out/r2c11_probe_mg_n$N.err
echo rc=$?; python - <<'PY'
import json
for line in open('gpurun_out/r2c11_probe_mg_n8.log'):
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d['spec'], 'ms', round(d['ms_per_step_max'],3), d['max_over_ranks'])
PY
tail -3 gpurun_out/r2c11_probe_mg_n$N.err
