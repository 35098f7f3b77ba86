#!/bin/bash
mkdir -p gpurun_out
cp sharkmer_b200/libsharkmer_b200.so /tmp/base.so
run() { # name lib ctas
  cp $2 sharkmer_b200/libsharkmer_b200.so
  SKM_INSERT_CTAS=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/m_$1.json 2> gpurun_out/m_$1.err || echo "$1 failed"
  python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/m_{n}.json')); s=d['stage_ms']
    print('%-12s value %.2f G/s step %.2f ms | e2e %.2f ms | ins %.2f part %.2f' % (n, d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], s['insert'], s['partition']))
except Exception as e: print(n,'ERR',e)
PY
}
run r32_c6 variants/lib_m7.so 6
run r32_c5 variants/lib_m7.so 5
run r40_c5 /tmp/base.so 5
cp /tmp/base.so sharkmer_b200/libsharkmer_b200.so
