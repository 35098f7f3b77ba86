#!/bin/bash
# quick check: count-parity tests + the default bench (no CPU leg)
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "count_parity or skewed or growth or reset or chunk_invariance or saturation" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
SKM_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; grep -m1 skm gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value']/1e9, 'e2e', d.get('e2e',{}).get('ms_per_step'))
print(d['stage_ms'])
PY
