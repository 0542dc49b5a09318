#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
N=${1:-2}; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_mg.py c4 "$@" > gpurun_out/r2_probe_probe_mg_n$N.log 2> gpurun_out/r2_probe_probe_mg_n$N.err
python - <<PY
import json
for line in open('gpurun_out/r2_probe_probe_mg_n$N.log'):
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d['spec'], 'ms', round(d['ms_per_step_max'],3), d['max_over_ranks'])
PY
