#!/bin/bash
# N=8 on the final build: C2 (weak, the headline shape) with a timeline, then C3 (strong, k=31, 100 M reads)
mkdir -p gpurun_out
TAG=r2_55
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/${TAG}_gpus.txt
run() {
  T=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --no-gups "$@" > gpurun_out/${T}.json 2> gpurun_out/${T}.err || { echo "$T FAILED"; tail -15 gpurun_out/${T}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}.json'))
print('${T}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'rounds', d.get('rounds'))
p=d.get('parity') or {}
print('   parity', {k:v for k,v in p.items() if k!='full_size_run'}, {k:v for k,v in (p.get('full_size_run') or {}).items() if k!='note'})
PY
}
SKM_TRACE=gpurun_out/${TAG}_trace_c2 run ${TAG}_c2_n8 --steps 5 --warmup 3
run ${TAG}_c3_n8 --config C3 --steps 3 --warmup 2 --no-e2e
