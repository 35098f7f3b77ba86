#!/bin/bash
# Final evidence for round 1: default bench (with CPU baseline + gups), launch list, full capture of
# the dominant kernel, reference arm.
mkdir -p gpurun_out
timeout 900 python bench.py --gups > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"insert_runs" -s 35 -c 2 -o gpurun_out/prof_r01b_insert $CMD > gpurun_out/ncu_insert.log 2>&1
echo "insert capture rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','bases_per_sec','stage_ms','roofline','e2e','cpu_baseline','gups','clocks','gpu_launches')}, indent=1)[:3500])
PY
