#!/bin/bash
# ncu evidence for the default bench command (N=1).  Launch list first, then a full capture of the
# top kernel (insert_runs_kernel) and of the other stages; each after the plain run exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"insert_runs|bucket_scatter|bucket_count|pack_kernel|histogram_kernel" -s 60 -c 10 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
# direct-mode top kernel for comparison
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --mode direct"
$CMD2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"extract_insert" -s 6 -c 2 -o gpurun_out/prof_r01_direct $CMD2 > gpurun_out/ncu_full_direct.log 2>&1
echo "direct capture rc=$?"
ls -la gpurun_out | grep -E "prof_r01|launches_r01"
