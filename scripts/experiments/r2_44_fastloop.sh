#!/bin/bash
# tile_insert_kernel: probe loop without a probe counter while the partition has room, 32-bit shared-window addresses
mkdir -p gpurun_out
TAG=${TAG:-r2_44}
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "count_parity or skewed or growth or reset or chunk_invariance or saturation or cluster_tile or sharded_group_vs or memory_bounded or c1_full" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-gups --no-services > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['steps_ms_wall'])
print(d['parity'])
PY
