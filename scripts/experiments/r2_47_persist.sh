#!/bin/bash
# persistent clusters in the cluster tile sort
mkdir -p gpurun_out
TAG=r2_47
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cluster_tile or sharded or count_parity_vs_oracle" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
for TL in 14 16; do
  SKM_TILE_LOG2=$TL timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-gups --no-services > gpurun_out/${TAG}_tl${TL}.json 2> gpurun_out/${TAG}_tl${TL}.err
  echo "bench tl=$TL exit $?"; tail -2 gpurun_out/${TAG}_tl${TL}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_tl${TL}.json'))
    print('tile_log2 ${TL}: ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()})
except Exception as e:
    print('no result', e)
PY
done
