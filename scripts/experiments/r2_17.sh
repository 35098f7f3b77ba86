#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -q -m gpu -x -k "sharded or group" > gpurun_out/r2_17_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_17_pytest.log
