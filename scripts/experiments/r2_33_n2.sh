#!/bin/bash
# N=2 after batching the owner kernels (one launch for all owners, warp-aggregated ranks) and merging the finalize all-gathers
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_33_gpus.txt
run() {
  TAG=$1; N=$2; shift; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-gups --no-services "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -12 gpurun_out/${TAG}.err; return; }
  else
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -12 gpurun_out/${TAG}.err; return; }
  fi
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'rounds', d.get('rounds'))
p=d.get('parity') or {}
print('   parity', {k:v for k,v in p.items() if k!='full_size_run'}, {k:v for k,v in (p.get('full_size_run') or {}).items() if k!='note'})
PY
}
run r2_33_c2_n1 1 --steps 5 --warmup 3
SKM_TRACE=gpurun_out/r2_33_trace_n8 run r2_33_c2_n8 8 --steps 5 --warmup 3
run r2_33_c2_n4 4 --steps 5 --warmup 3 --no-cpu
SKM_TRACE=gpurun_out/r2_33_trace_n2 run r2_33_c2_n2 2 --steps 5 --warmup 3 --no-cpu
