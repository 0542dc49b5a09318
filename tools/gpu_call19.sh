#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fa2 --converged-frac 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), d['phase_ms'], d['verify']['ok'], d['verify']['max_rel_err'])"; }
run base X=1
run overlap3 SVI_LS_OVERLAP_REFRESH=1
run overlap2 SVI_LS_OVERLAP_REFRESH=1 SVI_LS_S3_BLOCKS_PER_SM=2
run s3_2blocks_only SVI_LS_S3_BLOCKS_PER_SM=2
SVI_LS_OVERLAP_REFRESH=1 SVI_LS_S3_BLOCKS_PER_SM=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
