#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_insert' -c 2 -o gpurun_out/r2_09_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-gups > gpurun_out/r2_09_ncu.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/r2_09_ncu.log
