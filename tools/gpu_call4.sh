#!/bin/bash
# round 2, call 4 (2 GPUs): peer-memory exchange under torchrun vs the NCCL baseline; checksum vs N=1; CLI -gpus 2
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --checksum 6 > gpurun_out/r2c4_bench_c4_n2_peer.json 2> gpurun_out/r2c4_bench_c4_n2_peer.err; echo "peer rc=$?"
tail -3 gpurun_out/r2c4_bench_c4_n2_peer.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c4_bench_c4_n2_peer.json'))
    print({k:d[k] for k in ('value','ms_per_step','phase_ms','mg_phase_ms','verify','checksum','setup_s')}, d['e2e']['value'])
except Exception as e: print("parse failed", e)
PY
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl --no-verify > gpurun_out/r2c4_bench_c4_n2_nccl.json 2> gpurun_out/r2c4_bench_c4_n2_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c4_bench_c4_n2_nccl.json'))
    print({k:d[k] for k in ('value','ms_per_step','phase_ms')}, d['e2e']['value'])
except Exception as e: print("parse failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_cli.py -k gpus_n -x -q 2>&1 | tail -3
