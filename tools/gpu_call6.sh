#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=10
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_mg.py c4 \
   push=ce,chunks=4,ratio=1.0 push=ce,chunks=1 push=sm,chunks=1 push=sm,chunks=4,ratio=1.0 push=sm,chunks=3,ratio=0.7 push=ce_multi,chunks=3,ratio=0.7 push=sm,chunks=3,ratio=0.7,blocks=64 \
   > gpurun_out/r2c6_probe_mg_n$N.log 2> gpurun_out/r2c6_probe_mg_n$N.err
echo rc=$?; cat gpurun_out/r2c6_probe_mg_n$N.log; tail -5 gpurun_out/r2c6_probe_mg_n$N.err
