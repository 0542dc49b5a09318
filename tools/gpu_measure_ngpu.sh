#!/bin/bash
# final multi-GPU measurements: bench under torchrun at N GPUs (+ at 8: probe of a few exchange configs, CLI -gpus 8)
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --checksum 6 > gpurun_out/r2_ngpu_bench_n$N.json 2> gpurun_out/r2_ngpu_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_ngpu_bench_n$N.json').read().strip().split('\n')[-1])
    print('N=$N', {k:d[k] for k in ('value','ms_per_step','mg_phase_ms','checksum','setup_s')}, 'e2e', d['e2e']['value'], 'verify', d['verify'] and (d['verify']['ok'], d['verify']['max_rel_err']))
except Exception as e: print("parse failed", e)
PY
if [ "$N" = "8" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/probe_mg.py c4 \
     push=ce,chunks=4,ratio=0.6 push=ce,chunks=5,ratio=0.6 push=ce,chunks=4,ratio=0.7 push=ce,chunks=3,ratio=0.6 push=ce,chunks=6,ratio=0.7 \
     > gpurun_out/r2_ngpu_probe_mg_n8.log 2> gpurun_out/r2_ngpu_probe_mg_n8.err
  python - <<'PY'
import json
for line in open('gpurun_out/r2_ngpu_probe_mg_n8.log'):
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d['spec'], 'ms', round(d['ms_per_step_max'],3), d['max_over_ranks'])
PY
  timeout 900 bash tools/cli_gpus_compare.sh c3 8 8 > gpurun_out/r2_ngpu_cli_gpus8_c3.log 2>&1; echo "cli compare rc=$?"; tail -6 gpurun_out/r2_ngpu_cli_gpus8_c3.log
fi
