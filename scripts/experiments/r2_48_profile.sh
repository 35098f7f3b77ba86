#!/bin/bash
# profiles of the current build: launch list of the default bench + --set full of the kernels of the tiled path,
# and of the cluster tile sort (SKM_TILE_LOG2=16, what an 8-GPU sender runs)
mkdir -p gpurun_out
TAG=r2_48
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gups --no-services"
timeout 600 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_insert|tile_sort|bucket_scatter|pack_kernel' -s 40 -c 8 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full exit $?"
SKM_TILE_LOG2=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_sort_cluster' -s 12 -c 2 -o gpurun_out/${TAG}_prof_cluster $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "cluster exit $?"
SKM_TRACE=gpurun_out/${TAG}_timeline.txt timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-gups --no-services > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
echo "trace exit $?"
wc -l gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_timeline.txt; ls -la gpurun_out/${TAG}_prof*.ncu-rep
