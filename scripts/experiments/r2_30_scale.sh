#!/bin/bash
# final scaling lines on an 8-GPU box: C2 (weak) at N = 1, 2, 4, 8; C3 (strong, 100 M reads) at N = 8, 4, 2; C5 (1 B reads) at N = 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_30_gpus.txt
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
run r2_30_c2_n1 1 --steps 5 --warmup 3
SKM_TRACE=gpurun_out/r2_30_trace_n8 run r2_30_c2_n8 8 --steps 5 --warmup 3
run r2_30_c2_n4 4 --steps 5 --warmup 3 --no-cpu
run r2_30_c2_n2 2 --steps 5 --warmup 3 --no-cpu
run r2_30_c3_n8 8 --config C3 --steps 2 --warmup 1 --no-cpu
run r2_30_c3_n4 4 --config C3 --steps 2 --warmup 1 --no-cpu
run r2_30_c3_n2 2 --config C3 --steps 1 --warmup 1 --no-cpu --no-e2e
run r2_30_c5_n8 8 --config C5 --steps 1 --warmup 1 --no-cpu --no-e2e
