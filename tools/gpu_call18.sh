#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 1200 python -m pytest tests/test_gpu_mg.py tests/test_gpu_parity_ring.py tests/test_gpu_parity.py -x -q > gpurun_out/r2c18_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c18_tests.log
tail -4 gpurun_out/r2c18_tests.log
timeout 300 python tools/probe_late.py c4 2>&1 | tail -3
