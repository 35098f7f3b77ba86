#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_sharded_group_vs_oracle and 2-25" > gpurun_out/r2_18_san.log 2>&1
echo "exit $?"
grep -E "Invalid|at 0x|by thread|Address|in .*kernel|=========     in" gpurun_out/r2_18_san.log | head -30
