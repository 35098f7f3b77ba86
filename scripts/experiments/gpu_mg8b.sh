#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" > gpurun_out/lscpu.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for bind in 1 0; do
SKM_NUMA_BIND=$bind timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/mg${N}_bind$bind.json 2> gpurun_out/mg${N}_bind$bind.err; echo "bind$bind rc=$?"
grep -E "^\[rank" gpurun_out/mg${N}_bind$bind.err | sort | head -20
python - $N $bind <<'PY'
import json,os,sys
N,b=sys.argv[1:3]
f=f'mg{N}_bind{b}'
if os.path.exists(f'gpurun_out/{f}.json') and os.path.getsize(f'gpurun_out/{f}.json'):
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert']))
else: print(f,'FAILED')
PY
done
