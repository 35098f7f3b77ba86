#!/bin/bash
mkdir -p gpurun_out
for C in 4 5; do
SKM_INSERT_CTAS=$C timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c$C.json 2> gpurun_out/c$C.err; echo "c$C rc=$?"
python - $C <<'PY'
import json,sys
C=sys.argv[1]
d=json.load(open(f'gpurun_out/c{C}.json')); s=d['stage_ms']
print('ctas=%s value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f part %.2f' % (C, d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['partition']))
PY
done
