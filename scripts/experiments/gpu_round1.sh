#!/bin/bash
# First GPU pass: parity tests, smoke, first bench lines.  Run under gpurun from the repo root.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --mode direct --gups > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench direct rc=$?"; tail -c 3000 gpurun_out/bench_direct.json; tail -5 gpurun_out/bench_direct.err
timeout 600 python bench.py --steps 3 --warmup 3 --mode partitioned --no-cpu > gpurun_out/bench_part.json 2> gpurun_out/bench_part.err; echo "bench part rc=$?"; tail -c 3000 gpurun_out/bench_part.json; tail -5 gpurun_out/bench_part.err
