#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_13_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_13_pytest.log
run() {
  TAG=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 2 --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
print('   parity', d.get('parity'))
PY
}
run r2_13_c2
run r2_13_c1 --config C1 --no-cpu
