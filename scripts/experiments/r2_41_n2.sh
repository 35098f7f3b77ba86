#!/bin/bash
# N=2: slices (cluster-sorted sender list, no re-bucketing pass) vs the owner-list path
mkdir -p gpurun_out
TAG=r2_41
run() {
  T=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-gups --steps 5 --warmup 3 "$@" > gpurun_out/${T}.json 2> gpurun_out/${T}.err || { echo "$T FAILED"; tail -12 gpurun_out/${T}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}.json'))
print('${T}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
p=d.get('parity') or {}
print('   parity', {k:v for k,v in p.items() if k!='full_size_run'}, {k:v for k,v in (p.get('full_size_run') or {}).items() if k!='note'})
PY
}
SKM_TRACE=gpurun_out/${TAG}_trace_slices run ${TAG}_slices
SKM_MG_SLICES=0 run ${TAG}_lists --no-cpu
SKM_TILE_LOG2=15 run ${TAG}_slices_tl15 --no-cpu --no-e2e
