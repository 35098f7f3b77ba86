#!/bin/bash
# does the clock sampler perturb the steps?  nvidia-smi -lms 100 (recipe) vs in-process NVML vs none; per-step wall times
mkdir -p gpurun_out
TAG=r2_43
for S in smi nvml none smi2 nvml2; do
  EXTRA=""
  case $S in
    smi|smi2) export SKM_SAMPLER=smi; unset SKM_NO_SAMPLER;;
    nvml|nvml2) export SKM_SAMPLER=nvml; unset SKM_NO_SAMPLER;;
    none) export SKM_NO_SAMPLER=1;;
  esac
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-gups --no-services > gpurun_out/${TAG}_${S}.json 2> gpurun_out/${TAG}_${S}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_${S}.json'))
print('${S}: dev %.2f ms' % d['ms_per_step'], d['steps_ms_wall'], '| e2e %.2f ms' % d['e2e']['ms_per_step'], d['e2e']['steps_ms_wall'], d['clocks'])
PY
done
