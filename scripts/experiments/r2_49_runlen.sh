#!/bin/bash
# how sensitive is the insert kernel to the run length?  fewer pass-A buckets on one GPU = more partitions per bucket =
# shorter runs per tile (1024 buckets: 64 k-mers; 512: 32; 256: 16; 128: 8) -- what an N-GPU owner sees with small tiles
mkdir -p gpurun_out
TAG=r2_49
for MB in 512 256 128; do
  SKM_MAX_BUCKETS=$MB timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-gups --no-services > gpurun_out/${TAG}_mb${MB}.json 2> gpurun_out/${TAG}_mb${MB}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_mb${MB}.json'))
    print('max_buckets ${MB}: ms/step %.2f value %.2f G' % (d['ms_per_step'], d['value']/1e9), {k: round(v,2) for k,v in d['stage_ms'].items()})
except Exception as e:
    print('no result', e); import subprocess; print(open('gpurun_out/${TAG}_mb${MB}.err').read()[-600:])
PY
done
# the dominant kernel's full capture for profiles/
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-gups --no-services"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_insert' -s 3 -c 1 -o gpurun_out/${TAG}_prof_insert $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
