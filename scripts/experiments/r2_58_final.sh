#!/bin/bash
# final check of the round: whole GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
TAG=r2_58
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -1 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err
echo "reference exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['steps_ms_wall'], d['steps_ms_wall'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','streaming_achieved','streaming_frac','traffic','traffic_GBps','random_access_frac')})
print(d['clocks']); print(d['parity']['full_size_run']['table_digest_equals_oracle_golden'])
r=json.load(open('gpurun_out/${TAG}_reference.json')); print('reference', r['value'], r['cpu_baseline']['cores'])
PY
