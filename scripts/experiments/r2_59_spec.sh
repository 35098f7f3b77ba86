#!/bin/bash
# speculative insert launch (queued behind the lists' events, overflow checked on the device): whole GPU suite + bench
mkdir -p gpurun_out
TAG=r2_59
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -1 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['steps_ms_wall'], d['steps_ms_wall'])
print(d['parity'])
PY
