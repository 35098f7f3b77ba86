#!/bin/bash
# usage: gpu_mg.sh N   (N ranks)
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --exchange p2p > gpurun_out/mg${N}_p2p.json 2> gpurun_out/mg${N}_p2p.err; echo "p2p rc=$?"; tail -4 gpurun_out/mg${N}_p2p.err
timeout 900 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --exchange nccl > gpurun_out/mg${N}_nccl.json 2> gpurun_out/mg${N}_nccl.err; echo "nccl rc=$?"; tail -2 gpurun_out/mg${N}_nccl.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg1_full.json 2> gpurun_out/mg1_full.err; echo "n1 rc=$?"
python - $N <<'PY'
import json,os,sys
N=sys.argv[1]
for f in ('mg1_full',f'mg{N}_p2p',f'mg{N}_nccl'):
    if not os.path.exists(f'gpurun_out/{f}.json') or os.path.getsize(f'gpurun_out/{f}.json')==0:
        print(f,'FAILED'); continue
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f pack %.2f nvlink/step %.2f GB' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], s['pack'], d.get('nvlink_bytes_sent_per_step_rank0',0)/1e9))
PY
