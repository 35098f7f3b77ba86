#!/bin/bash
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for C in 4 5; do for X in p2p dma; do
SKM_INSERT_CTAS=$C timeout 900 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e --exchange $X > gpurun_out/mg${N}_${X}_c$C.json 2> gpurun_out/mg${N}_${X}_c$C.err; echo "$X ctas$C rc=$?"
done; done
python - $N <<'PY'
import json,os,sys
N=sys.argv[1]
for C in (4,5):
  for X in ('p2p','dma'):
    f=f'mg{N}_{X}_c{C}'
    if not os.path.exists(f'gpurun_out/{f}.json') or os.path.getsize(f'gpurun_out/{f}.json')==0:
        print(f,'FAILED'); continue
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | ins %.2f cnt %.2f part %.2f' % (d['value']/1e9, d['ms_per_step'], s['insert'], s['count'], s['partition']))
PY
