#!/bin/bash
# one copy stream per destination rank: sharded parity tests on one GPU (up to 8 ranks sharing it)
mkdir -p gpurun_out
TAG=r2_56
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -q -m gpu -x -k "sharded or gpus" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
