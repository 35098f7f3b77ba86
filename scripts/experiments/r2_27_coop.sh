#!/bin/bash
# A/B: warp-cooperative probing (8 lanes per k-mer) vs one thread per k-mer, direct kernel, loads 0.53 and ~0.7
mkdir -p gpurun_out
SKM_WARP_COOP=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "count_parity or growth or goldens or empty_reads or skewed" > gpurun_out/r2_27_pytest.log 2>&1
echo "pytest (coop) exit $?"; tail -3 gpurun_out/r2_27_pytest.log
run() {
  TAG=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-gups --no-e2e --no-services --mode direct "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || { echo "$TAG FAILED"; tail -5 gpurun_out/${TAG}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}.json'))
print('${TAG}', 'ms/step %.2f' % d['ms_per_step'], 'insert %.2f' % d['stage_ms']['insert'], 'load %.3f' % d['table']['load'])
PY
}
run r2_27_thread_k21
SKM_WARP_COOP=1 run r2_27_coop_k21
run r2_27_thread_k31 --k 31
SKM_WARP_COOP=1 run r2_27_coop_k31 --k 31
# the default (tiled) path with the table services leg and the GUPS probe, for the record
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_27_default.json 2> gpurun_out/r2_27_default.err; echo "default exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_27_default.json'))
print('default ms/step %.2f value %.2f G e2e %.2f G' % (d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9))
print(json.dumps(d.get('services'), indent=1)[:900])
print({k: (round(v,3) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('kernels','traffic_note','algorithmic_bytes','step_survey_model','random_access_note')})
print(d['roofline']['step_survey_model'])
for k,v in d['roofline']['kernels'].items(): print(' ', k, round(v['ms_per_step'],2), 'ms', round(v['achieved'],0), 'GB/s', round(v['frac'],3))
print(d['cpu_baseline'])
PY
