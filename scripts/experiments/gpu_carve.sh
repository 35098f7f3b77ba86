#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/trace_carve.txt
SKM_TRACE=gpurun_out/trace_carve.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/carve.json 2> gpurun_out/carve.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/carve.json')); s=d['stage_ms']
print('carveout: value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f part %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['partition']))
rows=[l.split() for l in open('gpurun_out/trace_carve.txt')]
bs=sorted(set(int(r[0]) for r in rows))
for b in bs[::-1]:
    rr=[r for r in rows if int(r[0])==b]
    if sum(1 for r in rr if r[1]=='insert')>=10 and not any(r[1]=='h2d' for r in rr):
        for r in rr: print('  %-10s %8.3f %8.3f  (%.3f)'%(r[1],float(r[2]),float(r[3]),float(r[3])-float(r[2])))
        break
PY
