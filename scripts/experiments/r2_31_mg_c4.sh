#!/bin/bash
# round 2: batched owner kernels + merged all-gathers (in-process ranks on one GPU), C4 at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "sharded or c4" > gpurun_out/r2_31_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_31_tests.log
tail -5 gpurun_out/r2_31_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_31_c2_n1.json 2> gpurun_out/r2_31_c2_n1.err
echo "bench rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/r2_31_c2_n1.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["stage_ms"])
PY
