#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
for R in 16 8 4; do
SKM_TILE_ROWS=$R timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/t_$R.json 2> gpurun_out/t_$R.err
python - $R <<'PY'
import json,sys
R=sys.argv[1]
d=json.load(open(f'gpurun_out/t_{R}.json')); s=d['stage_ms']
print('tile rows %s: value %.2f G/s step %.2f ms ins %.2f part %.2f clocks %s' % (R, d['value']/1e9, d['ms_per_step'], s['insert'], s['partition'], d['clocks']))
PY
done
