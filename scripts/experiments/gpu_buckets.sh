#!/bin/bash
mkdir -p gpurun_out
for NB in 1024 256 128 64; do
SKM_MAX_BUCKETS=$NB timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/nb$NB.json 2> gpurun_out/nb$NB.err; echo "nb$NB rc=$?"
python - $NB <<'PY'
import json,sys
C=sys.argv[1]
d=json.load(open(f'gpurun_out/nb{C}.json')); s=d['stage_ms']
print('buckets=%s value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f part %.2f' % (C, d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['partition']))
PY
done
