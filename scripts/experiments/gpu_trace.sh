#!/bin/bash
# timelines (SKM_TRACE) of one device-resident step and one host-buffer step, N=1
mkdir -p gpurun_out
rm -f gpurun_out/trace_dev.txt gpurun_out/trace_e2e.txt
SKM_TRACE=gpurun_out/trace_dev.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/tr_dev.json 2> gpurun_out/tr_dev.err; echo "dev rc=$?"
SKM_TRACE=gpurun_out/trace_all.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/tr_e2e.json 2> gpurun_out/tr_e2e.err; echo "e2e rc=$?"
python - <<'PY'
import json
for n in ('tr_dev','tr_e2e'):
    d=json.load(open(f'gpurun_out/{n}.json'))
    print(n, 'step %.2f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'))
def show(path):
    rows=[l.split() for l in open(path)]
    last=max(int(r[0]) for r in rows)
    for b in (last,):
        print(path,'batch',b)
        for r in rows:
            if int(r[0])==b: print('  %-10s %8.3f %8.3f  (%.3f)'%(r[1],float(r[2]),float(r[3]),float(r[3])-float(r[2])))
show('gpurun_out/trace_dev.txt'); show('gpurun_out/trace_all.txt')
PY
