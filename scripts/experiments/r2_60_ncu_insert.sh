#!/bin/bash
# --set full of the final tile_insert_kernel (the same command ran plain in r2_59)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tile_insert' -s 3 -c 1 -o gpurun_out/r2_60_prof_insert python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-gups --no-services > gpurun_out/r2_60_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2_60_prof_insert.ncu-rep
