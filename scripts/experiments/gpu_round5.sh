#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu"
run() {
  name=$1; shift
  env $ENVV $B "$@" > gpurun_out/z_$name.json 2> gpurun_out/z_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/z_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/z_{n}.json'))
s=d['stage_ms']
print(f"{n:28s} value={d['value']/1e9:7.2f} G/s step={d['ms_per_step']:7.2f} ms insert={s['insert']:7.2f} hist={s['histogram']:6.2f} count={s['count']:5.2f} part={s['partition']:6.2f} pack={s['pack']:5.2f} load={d['table']['load']:.2f} slots=2^{d['table']['slots'].bit_length()-1}")
PY
}
for R in 17 18 19 20; do
  ENVV="SKM_REGION_LOG2=$R" run part_c10_r$R --mode partitioned
done
ENVV="SKM_REGION_LOG2=17" run part_c1_r17 --mode partitioned --chunks 1
ENVV="SKM_REGION_LOG2=19" run part_c1_r19 --mode partitioned --chunks 1
ENVV="" run part_k31_c1 --mode partitioned --chunks 1 --k 31
ENVV="" run direct_k31_c1 --mode direct --chunks 1 --k 31
