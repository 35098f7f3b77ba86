#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --mode partitioned"
for R in 19 21 22 23 24 25; do
  SKM_REGION_LOG2=$R $B > gpurun_out/r_$R.json 2> gpurun_out/r_$R.err && python -c "
import json
d=json.load(open('gpurun_out/r_$R.json')); s=d['stage_ms']
print('region 2^$R slots (%d MB): step %.2f insert %.2f count %.2f part %.2f' % (16*2**$R/2**20, d['ms_per_step'], s['insert'], s['count'], s['partition']))"
done
