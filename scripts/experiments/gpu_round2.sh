#!/bin/bash
# tests + gups variants + ncu on the two insert kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --mode direct --gups --no-e2e --no-cpu > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench direct rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_direct.json')); print(json.dumps(d['gups'],indent=1)); print(d['stage_ms'])"
# ncu: small workload so it finishes fast (2M reads), after the plain run of the same command
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --reads-per-gpu 2000000"
$CMD --mode direct > gpurun_out/plain_direct.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extract_insert -s 12 -c 2 -o gpurun_out/prof_direct $CMD --mode direct > gpurun_out/ncu_direct.log 2>&1
echo "ncu direct rc=$?"
$CMD --mode partitioned > gpurun_out/plain_part.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"insert_sorted|bucket_scatter|bucket_count" -s 30 -c 6 -o gpurun_out/prof_part $CMD --mode partitioned > gpurun_out/ncu_part.log 2>&1
echo "ncu part rc=$?"
ls -la gpurun_out
