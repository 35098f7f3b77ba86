#!/bin/bash
# pack kernels on their own stream (pass A of batch b+1 overlaps pass B of batch b); cluster sort keeps 16-bit tile cells
mkdir -p gpurun_out
TAG=r2_42
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cluster_tile or sharded_group_vs or count_parity_vs_oracle or skewed or reset or ingest_reads" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
for TL in 13 16; do
  SKM_TRACE=gpurun_out/${TAG}_trace_tl${TL}.txt SKM_TILE_LOG2=$TL timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-gups --no-services > gpurun_out/${TAG}_tl${TL}.json 2> gpurun_out/${TAG}_tl${TL}.err
  echo "bench tl=$TL exit $?"; tail -2 gpurun_out/${TAG}_tl${TL}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_tl${TL}.json'))
    print('tile_log2 ${TL}: ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['stage_ms'])
except Exception as e:
    print('no result', e)
PY
done
