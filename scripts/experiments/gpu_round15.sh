#!/bin/bash
mkdir -p gpurun_out
for R in 16 8 4 2 1; do
SKM_TILE_ROWS=$R timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/t_$R.json 2> gpurun_out/t_$R.err; echo "rows $R rc=$?"
done
python - <<'PY'
import json,os
for R in (16,8,4,2,1):
    f=f'gpurun_out/t_{R}.json'
    if not os.path.exists(f) or os.path.getsize(f)==0: print(R,'FAILED'); continue
    d=json.load(open(f)); s=d['stage_ms']
    print('tile rows %2d: value %.2f G/s step %.2f ms ins %.2f' % (R, d['value']/1e9, d['ms_per_step'], s['insert']))
PY
