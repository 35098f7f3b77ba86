#!/bin/bash
N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
rm -f gpurun_out/trace_mg2.txt.rank*
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SKM_TRACE=gpurun_out/trace_mg2.txt timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/mg2_tr.json 2> gpurun_out/mg2_tr.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/mg2_tr.json')); s=d['stage_ms']
print('N=2 value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition']))
rows=[l.split() for l in open('gpurun_out/trace_mg2.txt.rank0')]
bs=sorted(set(int(r[0]) for r in rows))
for b in bs[::-1]:
    rr=[r for r in rows if int(r[0])==b]
    if sum(1 for r in rr if r[1]=='insert')>=10 and not any(r[1]=='h2d' for r in rr):
        for r in rr: print('  %-10s %8.3f %8.3f  (%.3f)'%(r[1],float(r[2]),float(r[3]),float(r[3])-float(r[2])))
        break
PY
