#!/bin/bash
# round 2, call 7: full GPU test suite (new multi-rank path included), bench N=1, insert-kernel variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2_07_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_07_pytest.log
run() {
  TAG=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 2 "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
}
run r2_07_t512c2
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=256 -DSKM_INS_CTAS=3" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_07_t256c3 --no-cpu --no-e2e --no-gups
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=256 -DSKM_INS_CTAS=4" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_07_t256c4 --no-cpu --no-e2e --no-gups
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=1024 -DSKM_INS_CTAS=1" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_07_t1024c1 --no-cpu --no-e2e --no-gups
