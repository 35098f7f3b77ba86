#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for C in 3 4 5 6; do
SKM_INSERT_CTAS=$C timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/n1_ctas$C.json 2> gpurun_out/n1_ctas$C.err; echo "ctas $C rc=$?"
done
SKM_EAGER=0 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/n1_noeager.json 2> gpurun_out/n1_noeager.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --chunks 1 > gpurun_out/n1_c1.json 2> gpurun_out/n1_c1.err
python - <<'PY'
import json,os
for f in ('n1_ctas3','n1_ctas4','n1_ctas5','n1_ctas6','n1_noeager','n1_c1'):
    if not os.path.exists(f'gpurun_out/{f}.json') or os.path.getsize(f'gpurun_out/{f}.json')==0:
        print(f,'FAILED'); continue
    d=json.load(open(f'gpurun_out/{f}.json')); s=d['stage_ms']
    print(f, 'value %.2f G/s step %.2f ms | e2e %.2f G/s %.2f ms | ins %.2f cnt %.2f part %.2f pack %.2f hist %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], s['insert'], s['count'], s['partition'], s['pack'], s['histogram']))
PY
