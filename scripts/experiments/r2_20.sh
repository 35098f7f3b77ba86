#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_20_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2_20_pytest.log
