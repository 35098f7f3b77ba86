#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg1_full.json 2> gpurun_out/mg1_full.err; echo "n1 rc=$?"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for X in dma; do
timeout 900 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --exchange $X > gpurun_out/mg${N}_${X}.json 2> gpurun_out/mg${N}_${X}.err; echo "$X rc=$?"; tail -3 gpurun_out/mg${N}_${X}.err
done
python - $N <<'PY'
import json,os,sys
N=sys.argv[1]
for f in ('mg1_full',f'mg{N}_dma'):
    if not os.path.exists(f'gpurun_out/{f}.json') or os.path.getsize(f'gpurun_out/{f}.json')==0:
        print(f,'FAILED'); continue
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f launches/step %d' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], d['roofline']['launches_per_step']))
PY
