#!/bin/bash
# round 2, call 2: first run of the tiled insert (tile_sort + tile_insert): parity tests, then the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_02_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2_02_pytest.log
tail -30 gpurun_out/r2_02_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_02_bench.json 2> gpurun_out/r2_02_bench.err
echo "bench exit $?"
tail -5 gpurun_out/r2_02_bench.err
head -c 1500 gpurun_out/r2_02_bench.json
