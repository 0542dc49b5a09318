#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
N=8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_mg.py c4 \
   push=ce,chunks=4,ratio=1.0 push=sm,chunks=4,ratio=1.0,blocks=32 push=sm,chunks=4,ratio=1.0,blocks=128 push=ce_multi,chunks=4,ratio=1.0 \
   push=ce_multi,chunks=4,ratio=0.6 push=sm,chunks=6,ratio=0.7,blocks=64 push=ce_multi,chunks=6,ratio=0.7 push=ce_multi,chunks=8,ratio=0.8 push=ce_multi,chunks=1 \
   > gpurun_out/r2c8_probe_mg_n$N.log 2> gpurun_out/r2c8_probe_mg_n$N.err
echo rc=$?; python - <<'PY'
import json
for line in open('gpurun_out/r2c8_probe_mg_n8.log'):
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d['spec'], 'ms', round(d['ms_per_step_max'],3), d['max_over_ranks'])
PY
tail -3 gpurun_out/r2c8_probe_mg_n$N.err
