#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 900 python -m pytest tests/test_gpu_mg.py tests/test_gpu_parity_ring.py tests/test_gpu_cli.py -x -q -k "not fa2" > gpurun_out/r2c7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c7_tests.log
tail -6 gpurun_out/r2c7_tests.log
