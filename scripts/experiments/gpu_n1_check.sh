#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/n1.json 2> gpurun_out/n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/n1.json')); s=d['stage_ms']
print('value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition']))
print(d['e2e']['stage_ms'])
PY
