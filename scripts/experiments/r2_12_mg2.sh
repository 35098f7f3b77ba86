#!/bin/bash
# 2 GPUs: the torchrun path of bench.py (IPC arenas, gloo all-gathers, NCCL barriers)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_12_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-gups > gpurun_out/r2_12_mg2.json 2> gpurun_out/r2_12_mg2.err
echo "exit $?"
tail -15 gpurun_out/r2_12_mg2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_12_mg2.json'))
print('ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
print(d.get('parity'))
PY
