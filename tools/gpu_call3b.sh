#!/bin/bash
mkdir -p gpurun_out
export SVI_LS_MG_TIMEOUT_S=5
timeout 900 python -m pytest tests/test_gpu_mg.py tests/test_gpu_cli.py -x -q > gpurun_out/r2c3b_mg_tests.log 2>&1; echo "mg+cli tests rc=$?" >> gpurun_out/r2c3b_mg_tests.log
tail -15 gpurun_out/r2c3b_mg_tests.log
