#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg1_full.json 2> gpurun_out/mg1_full.err; echo "n1 rc=$?"
SKM_EAGER=0 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg1_noeager.json 2> gpurun_out/mg1_noeager.err; echo "n1 noeager rc=$?"
if [ "$1" == "2" ]; then
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg2_full.json 2> gpurun_out/mg2_full.err; echo "full rc=$?"; tail -3 gpurun_out/mg2_full.err
fi
python - <<'PY'
import json,os
for f in ('mg1_full','mg1_noeager','mg2_full'):
    if not os.path.exists(f'gpurun_out/{f}.json') or os.path.getsize(f'gpurun_out/{f}.json')==0: continue
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f pack %.2f nvlink/step %.2f GB' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], s['pack'], d.get('nvlink_bytes_sent_per_step_rank0',0)/1e9))
PY
