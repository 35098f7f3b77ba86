#!/bin/bash
# 2 GPUs: C3 and C5 at reduced size through the sharded path (strong scaling configs, chunks = 0, flush rounds)
mkdir -p gpurun_out
run() {
  TAG=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -12 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()}, d['table'])
print('   parity', d.get('parity'))
PY
}
run r2_15_c3_quarter --config C3 --reads-per-gpu 25000000 --steps 2 --warmup 1
run r2_15_c5_16th --config C5 --reads-per-gpu 62500000 --steps 1 --warmup 1 --no-e2e
