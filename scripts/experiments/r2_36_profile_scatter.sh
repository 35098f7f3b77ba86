#!/bin/bash
# ncu --set full of the one-hash scatter kernel (after the same command ran plain)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-gups --no-services"
timeout 600 $CMD > gpurun_out/r2_36_plain.json 2> gpurun_out/r2_36_plain.err || { tail -5 gpurun_out/r2_36_plain.err; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bucket_scatter' -s 30 -c 2 -o gpurun_out/r2_36_prof $CMD > gpurun_out/r2_36_ncu.log 2>&1
echo "full exit $?"
