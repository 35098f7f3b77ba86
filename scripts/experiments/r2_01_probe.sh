#!/bin/bash
# round 2, call 1: shared-memory insert probe + baseline bench of the round-1 build on the same box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_01_gpu.txt
timeout 600 bench/smem_probe > gpurun_out/r2_01_probe.txt 2>&1
echo "probe exit $?" >> gpurun_out/r2_01_probe.txt
timeout 900 python bench.py --steps 3 --warmup 3 --gups --no-cpu > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/r2_01_probe.txt
