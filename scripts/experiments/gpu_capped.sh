#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for CAP in 1 0; do
rm -f gpurun_out/trace_cap$CAP.txt
SKM_CAPPED=$CAP SKM_TRACE=gpurun_out/trace_cap$CAP.txt timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/cap$CAP.json 2> gpurun_out/cap$CAP.err; echo "cap$CAP rc=$?"
python - $CAP <<'PY'
import json,sys
C=sys.argv[1]
d=json.load(open(f'gpurun_out/cap{C}.json')); s=d['stage_ms']
print('capped=%s value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f launches %d' % (C, d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], d['gpu_launches']))
PY
done
