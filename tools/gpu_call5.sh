#!/bin/bash
# round 2, call 5 (8 GPUs): peer-memory exchange at 8 and 4 GPUs
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
run() { # n tag extra...
  n=$1; tag=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r2c5_$tag.json 2> gpurun_out/r2c5_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c5_$tag.json').read().strip().split('\n')[-1])
    print('$tag', {k:d[k] for k in ('value','ms_per_step','mg_phase_ms','verify','checksum','setup_s')}, 'e2e', d['e2e']['value'])
except Exception as e: print("parse failed", e)
PY
}
run 8 n8_c4 --checksum 6
run 8 n8_c2 --chunks 2 --no-verify
run 4 n4_c4 --no-verify
