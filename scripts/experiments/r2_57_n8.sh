#!/bin/bash
# N=8, C2, after the per-destination copy streams
mkdir -p gpurun_out
TAG=r2_57
T=${TAG}_c2_n8
SKM_TRACE=gpurun_out/${TAG}_trace_c2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-gups --steps 5 --warmup 3 > gpurun_out/${T}.json 2> gpurun_out/${T}.err || { echo "$T FAILED"; tail -15 gpurun_out/${T}.err; exit 1; }
python - <<PY
import json
d=json.load(open('gpurun_out/${T}.json'))
print('${T}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
p=d.get('parity') or {}
print('   parity', {k:v for k,v in p.items() if k!='full_size_run'}, {k:v for k,v in (p.get('full_size_run') or {}).items() if k!='note'})
PY
