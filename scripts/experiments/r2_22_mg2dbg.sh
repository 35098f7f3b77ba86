#!/bin/bash
mkdir -p gpurun_out
SKM_BENCH_DEBUG=1 SKM_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-gups --no-cpu > gpurun_out/r2_22.json 2> gpurun_out/r2_22.err
echo "exit $?"
grep bench gpurun_out/r2_22.err | head -30
grep "rank 0 host" gpurun_out/r2_22.err
