#!/bin/bash
# round 2, call 3: parity tests of the tiled insert, then launch list + full ncu capture of tile_insert / tile_sort / scatter
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_03_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2_03_pytest.log
tail -40 gpurun_out/r2_03_pytest.log
SKM_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_03_bench.json 2> gpurun_out/r2_03_bench.err
echo "bench exit $?"
grep skm gpurun_out/r2_03_bench.err | head -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_03_bench.json'))
print(d['ms_per_step'], d['stage_ms'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_insert|tile_sort|bucket_scatter' -s 20 -c 5 -o gpurun_out/r2_03_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_03_ncu.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/r2_03_ncu.log
