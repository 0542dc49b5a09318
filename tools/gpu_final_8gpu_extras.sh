#!/bin/bash
# 8 GPUs, same box: the NCCL-between-phases baseline (sharded.py exchange="nccl") next to the peer-memory exchange, and
# the C++ CLI at config 4 with -gpus 8
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
for ex in nccl peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --exchange $ex --no-verify > gpurun_out/r2_8gpu_$ex.json 2> gpurun_out/r2_8gpu_$ex.err; echo "$ex rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_8gpu_$ex.json').read().strip().split('\n')[-1])
    print('$ex', d['value'], d['ms_per_step'], d['phase_ms'], 'e2e', d['e2e']['value'])
except Exception as e: print("parse failed", e)
PY
done
timeout 900 bash tools/cli_at_scale.sh c4 5 -gpus 8 > gpurun_out/r2_cli_c4_gpus8.log 2>&1; echo "cli rc=$?"; tail -25 gpurun_out/r2_cli_c4_gpus8.log
