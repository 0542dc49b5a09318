#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 300 python tools/probe_phi.py c4 8 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_mg.py tests/test_gpu_parity_ring.py -x -q > gpurun_out/r2c10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c10_tests.log
tail -4 gpurun_out/r2c10_tests.log
