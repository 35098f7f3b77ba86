#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "capped_layouts" > gpurun_out/r2_32_dbg.log 2>&1
grep -n "AssertionError\|passed\|failed" gpurun_out/r2_32_dbg.log | head
