#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2_10_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_10_pytest.log
run() {
  TAG=$1; shift
  SKM_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], {k: round(v,2) for k,v in d['stage_ms'].items()})
PY
  grep -m1 "tile_insert" gpurun_out/${TAG}.err
}
run r2_10_base
build() { SKM_NVCC_EXTRA="$1" python -m sharkmer_b200.build --force > /dev/null 2>&1; }
build "-DSKM_INS_THREADS=256 -DSKM_INS_CTAS=3 -DSKM_INS_STAGE=512 -DSKM_INS_DEPTH=3" && run r2_10_t256c3
build "-DSKM_INS_THREADS=384 -DSKM_INS_CTAS=3 -DSKM_INS_STAGE=512 -DSKM_INS_DEPTH=3" && run r2_10_t384c3
build "-DSKM_INS_THREADS=1024 -DSKM_INS_CTAS=2 -DSKM_INS_STAGE=1024 -DSKM_INS_DEPTH=4" && run r2_10_t1024c2
build "" 
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_insert' -c 1 -o gpurun_out/r2_10_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-gups > gpurun_out/r2_10_ncu.log 2>&1
echo "ncu exit $?"
