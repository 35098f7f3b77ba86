#!/bin/bash
N=2
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for C in 5 6; do
SKM_INSERT_CTAS=$C timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/mg2_c$C.json 2> gpurun_out/mg2_c$C.err; echo "c$C rc=$?"
python - $C <<'PY'
import json,sys
C=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/mg2_c{C}.json')); s=d['stage_ms']
    print('N=2 ctas %s: value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f' % (C, d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert']))
except Exception as e: print('ERR', e)
PY
done
