#!/bin/bash
# 2-GPU run: N=2 bench through torchrun (NCCL), small then full
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu --reads-per-gpu 1000000 > gpurun_out/mg2_small.json 2> gpurun_out/mg2_small.err; echo "small rc=$?"; tail -c 1500 gpurun_out/mg2_small.json; tail -5 gpurun_out/mg2_small.err
timeout 900 $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg2_full.json 2> gpurun_out/mg2_full.err; echo "full rc=$?"; tail -c 2500 gpurun_out/mg2_full.json; tail -5 gpurun_out/mg2_full.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/mg1_full.json 2> gpurun_out/mg1_full.err; echo "n1 rc=$?"; python -c "
import json
for f in ('mg1_full','mg2_full'):
    d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['value']/1e9, d['ms_per_step'], d.get('e2e',{}).get('value',0)/1e9)"
