#!/bin/bash
# the whole GPU suite on the current build, then the default bench line
mkdir -p gpurun_out
TAG=${TAG:-r2_52}
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['steps_ms_wall'])
print({k:v for k,v in d['roofline'].items() if k!='kernels'})
print(d.get('cpu_baseline')); print(d.get('clocks')); print(d.get('services'))
PY
