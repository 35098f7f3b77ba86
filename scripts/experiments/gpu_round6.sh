#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu"
run() {
  name=$1; shift
  env $ENVV $B "$@" > gpurun_out/w_$name.json 2> gpurun_out/w_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/w_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/w_{n}.json'))
s=d['stage_ms']
print(f"{n:28s} value={d['value']/1e9:7.2f} G/s step={d['ms_per_step']:7.2f} ms insert={s['insert']:7.2f} hist={s['histogram']:6.2f} count={s['count']:5.2f} part={s['partition']:6.2f} pack={s['pack']:5.2f} load={d['table']['load']:.2f} slots=2^{d['table']['slots'].bit_length()-1}")
PY
}
ENVV="" run part_c10 --mode partitioned
ENVV="SKM_REGION_LOG2=20" run part_c10_r20 --mode partitioned
ENVV="" run part_c1 --mode partitioned --chunks 1
ENVV="" run direct_c10 --mode direct
