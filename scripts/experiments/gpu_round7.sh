#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu"
run() {
  name=$1; shift
  env $ENVV $B "$@" > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/v_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/v_{n}.json'))
s=d['stage_ms']
e=d.get('e2e',{})
print(f"{n:20s} value={d['value']/1e9:6.2f} G/s step={d['ms_per_step']:6.2f} wall={d['ms_per_step_wall']:6.2f} ins={s['insert']:6.2f} hist={s['histogram']:5.2f} cnt={s['count']:5.2f} part={s['partition']:5.2f} pack={s['pack']:5.2f} | e2e={e.get('value',0)/1e9:6.2f} G/s {e.get('ms_per_step',0):6.2f} ms h2d={e.get('stage_ms',{}).get('h2d',0):5.2f}")
PY
}
ENVV="" run auto_c10 --mode auto
ENVV="" run part_c10 --mode partitioned
ENVV="" run direct_c10 --mode direct
ENVV="" run auto_c1 --mode auto --chunks 1
