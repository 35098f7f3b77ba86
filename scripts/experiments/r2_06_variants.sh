#!/bin/bash
# tile_insert variants: threads x CTAs per SM (rebuilds the library on the box)
mkdir -p gpurun_out
run() {
  TAG=$1
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-gups > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "count_parity or skewed or growth or reset or chunk_invariance" > gpurun_out/r2_06_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_06_pytest.log
run r2_06_t512c2
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=256 -DSKM_INS_CTAS=3" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_06_t256c3
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=256 -DSKM_INS_CTAS=4" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_06_t256c4
SKM_NVCC_EXTRA="-DSKM_INS_THREADS=1024 -DSKM_INS_CTAS=1" python -m sharkmer_b200.build --force > /dev/null 2>&1 && run r2_06_t1024c1
