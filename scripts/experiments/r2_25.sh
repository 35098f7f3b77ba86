#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "count_parity or skewed or growth or reset or sharded or memory or flush or c2_full" > gpurun_out/r2_25_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_25_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-gups > gpurun_out/r2_25_n1.json 2> gpurun_out/r2_25_n1.err; echo "n1 exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_25_n1.json'))
print('N=1 ms/step %.2f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
SKM_BENCH_DEBUG=1 SKM_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-gups --no-cpu > gpurun_out/r2_25_n2.json 2> gpurun_out/r2_25_n2.err
echo "n2 exit $?"
grep -E "rank 0 host|\[bench\] reset" gpurun_out/r2_25_n2.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_25_n2.json'))
print('N=2 ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
