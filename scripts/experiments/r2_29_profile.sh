#!/bin/bash
# final profiles of the default bench (C2, N=1): launch list + --set full of the three kernels of the tiled path
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gups --no-services"
timeout 600 $CMD > gpurun_out/r2_29_plain.json 2> gpurun_out/r2_29_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_29_launches.csv $CMD > gpurun_out/r2_29_ncu1.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_insert|tile_sort|bucket_scatter|pack_kernel' -s 60 -c 8 -o gpurun_out/r2_29_prof $CMD > gpurun_out/r2_29_ncu2.log 2>&1
echo "full exit $?"
SKM_TRACE=gpurun_out/r2_29_timeline.txt timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-gups --no-services > gpurun_out/r2_29_trace.json 2> gpurun_out/r2_29_trace.err
echo "trace exit $?"
wc -l gpurun_out/r2_29_launches.csv gpurun_out/r2_29_timeline.txt
