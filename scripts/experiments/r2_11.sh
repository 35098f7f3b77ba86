#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2_11_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_11_pytest.log
run() {
  TAG=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
}
run r2_11_overlap
SKM_SORT_OVERLAP=0 run r2_11_nooverlap
