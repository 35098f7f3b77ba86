#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_pcr.py -q -m gpu -x -k "scan or primer or pcr or lookup" > gpurun_out/r2_28_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_28_pytest.log
run() {
  TAG=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['ms_per_step'], d['clocks'], (d.get('services') or {}).get('scan_oligos'))
PY
}
run r2_28_s250
SKM_NO_SAMPLER=1 run r2_28_nosampler
SKM_SAMPLER_MS=1000 run r2_28_s1000
SKM_SAMPLER_MS=100 run r2_28_s100
