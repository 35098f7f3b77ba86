#!/bin/bash
# sort kernels compiled for 2 / 4 sub-buckets per thread: parity subset, then the one-GPU bench at tile 2^13 (default) and 2^16
mkdir -p gpurun_out
TAG=r2_54
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cluster_tile or sharded_group_vs or sharded_group_few or count_parity_vs_oracle or with_capacity_hint" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
for TL in 13 16; do
  SKM_TILE_LOG2=$TL timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-gups --no-services > gpurun_out/${TAG}_tl${TL}.json 2> gpurun_out/${TAG}_tl${TL}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_tl${TL}.json'))
    print('tile_log2 ${TL}: ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()})
except Exception as e:
    print('no result', e)
PY
done
