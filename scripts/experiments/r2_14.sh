#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_14_trace_c1.txt
SKM_TRACE=gpurun_out/r2_14_trace_c1.txt timeout 600 python bench.py --config C1 --steps 2 --warmup 1 --no-cpu --no-gups --no-e2e > gpurun_out/r2_14_c1.json 2> gpurun_out/r2_14_c1.err
echo "c1 exit $?"; tail -3 gpurun_out/r2_14_c1.err
tail -40 gpurun_out/r2_14_trace_c1.txt
run() {
  TAG=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 2 --no-gups "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'value %.2f G' % (d['value']/1e9), 'e2e', d.get('e2e',{}).get('ms_per_step'), {k: round(v,2) for k,v in d['stage_ms'].items()}, d['table'])
print('   parity', d.get('parity'))
PY
}
run r2_14_c3_eighth --config C3 --reads-per-gpu 12500000 --no-cpu
