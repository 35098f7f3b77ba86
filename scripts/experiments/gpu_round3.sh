#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu"
run() { # name, env...
  name=$1; shift
  env "$@" $B --mode ${MODE:-direct} > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/x_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/x_{n}.json'))
s=d['stage_ms']
print(f"{n:28s} value={d['value']/1e9:7.2f} G/s step={d['ms_per_step']:7.2f} ms insert={s['insert']:7.2f} hist={s['histogram']:6.2f} count={s['count']:5.2f} part={s['partition']:6.2f} pack={s['pack']:5.2f} load={d['table']['load']:.2f}")
PY
}
for D in 1 2 4 8; do
  run d${D}_track SKM_PIPE_DEPTH=$D
  run d${D}_scan SKM_PIPE_DEPTH=$D SKM_HISTO_SCAN=1
done
for F in 32 64 128; do
  run d4_fetch$F SKM_PIPE_DEPTH=4 SKM_L2_FETCH=$F
  run d8_fetch$F SKM_PIPE_DEPTH=8 SKM_L2_FETCH=$F
done
MODE=partitioned run part_d4 SKM_PIPE_DEPTH=4
MODE=partitioned run part_d8 SKM_PIPE_DEPTH=8
MODE=partitioned run part_d8_f32 SKM_PIPE_DEPTH=8 SKM_L2_FETCH=32
# gups under fetch granularities
for F in 32 64 128; do
  SKM_L2_FETCH=$F $B --gups --mode direct > gpurun_out/gups_f$F.json 2>/dev/null && python -c "
import json; d=json.load(open('gpurun_out/gups_f$F.json')); print('fetch $F', {k: round(v/1e9,1) for k,v in d['gups'].items()})"
done
