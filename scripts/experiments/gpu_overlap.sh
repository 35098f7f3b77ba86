#!/bin/bash
# A/B: insert kernel footprint (launch bounds, CTAs per SM) x stream priority, so that the bucketing
# kernels of later chunks co-run with the insert kernel
mkdir -p gpurun_out
cp sharkmer_b200/libsharkmer_b200.so /tmp/base.so
run() { # name lib ctas prio
  cp $2 sharkmer_b200/libsharkmer_b200.so
  SKM_INSERT_CTAS=$3 SKM_PRIO=$4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ov_$1.json 2> gpurun_out/ov_$1.err || echo "$1 failed"
  python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ov_{n}.json')); s=d['stage_ms']
    print('%-14s value %.2f G/s step %.2f ms | e2e %.2f ms | ins %.2f cnt %.2f part %.2f pack %.2f' % (n, d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], s['pack']))
except Exception as e: print(n,'ERR',e)
PY
}
run base_c6_p0 /tmp/base.so 6 0
run base_c6_p1 /tmp/base.so 6 1
run base_c4_p1 /tmp/base.so 4 1
run lb6_c6_p0 variants/lib_lb6.so 6 0
run lb6_c6_p1 variants/lib_lb6.so 6 1
run lb6_c5_p1 variants/lib_lb6.so 5 1
run lb6_c4_p1 variants/lib_lb6.so 4 1
run lb6_c3_p1 variants/lib_lb6.so 3 1
cp /tmp/base.so sharkmer_b200/libsharkmer_b200.so
