#!/bin/bash
mkdir -p gpurun_out
SKM_SYNC_DEBUG=1 SKM_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_sharded_group_vs_oracle and 2-25" > gpurun_out/r2_19_dbg.log 2>&1
echo "exit $?"
grep -E "SkmError|fault after|tile_insert:" gpurun_out/r2_19_dbg.log | head
