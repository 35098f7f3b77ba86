#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -q -m gpu -x -k "count_parity or skewed or growth or reset or sharded or memory or flush or cli or histo or paired" > gpurun_out/r2_26_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_26_pytest.log
run() {
  TAG=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'), d.get('e2e',{}).get('stage_ms'), d['table'])
PY
}
run r2_26_early
SKM_NO_EARLY_FLUSH=1 run r2_26_noearly
