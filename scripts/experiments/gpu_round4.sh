#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu"
run() { # name, extra args..., env via ENVV
  name=$1; shift
  env $ENVV $B "$@" > gpurun_out/y_$name.json 2> gpurun_out/y_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/y_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open(f'gpurun_out/y_{n}.json'))
s=d['stage_ms']
print(f"{n:28s} value={d['value']/1e9:7.2f} G/s step={d['ms_per_step']:7.2f} ms insert={s['insert']:7.2f} hist={s['histogram']:6.2f} count={s['count']:5.2f} part={s['partition']:6.2f} pack={s['pack']:5.2f} load={d['table']['load']:.2f} slots=2^{d['table']['slots'].bit_length()-1}")
PY
}
ENVV="SKM_PIPE_DEPTH=1" run part_c10_d1 --mode partitioned
ENVV="SKM_PIPE_DEPTH=2" run part_c10_d2 --mode partitioned
ENVV="SKM_PIPE_DEPTH=1" run part_c1_d1 --mode partitioned --chunks 1
ENVV="SKM_PIPE_DEPTH=2" run part_c1_d2 --mode partitioned --chunks 1
ENVV="SKM_PIPE_DEPTH=4" run part_c1_d4 --mode partitioned --chunks 1
ENVV="SKM_PIPE_DEPTH=1" run direct_c1_d1 --mode direct --chunks 1
ENVV="SKM_PIPE_DEPTH=1" run direct_c10_d1 --mode direct
ENVV="SKM_PIPE_DEPTH=1" run part_c0_d1 --mode partitioned --chunks 0
ENVV="SKM_PIPE_DEPTH=1" run direct_c0_d1 --mode direct --chunks 0
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --reads-per-gpu 4000000 --chunks 1 --mode partitioned"
$CMD > gpurun_out/plain_part1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"insert_list" -s 1 -c 1 -o gpurun_out/prof_part1 $CMD > gpurun_out/ncu_part1.log 2>&1
echo "ncu rc=$?"
