#!/bin/bash
# Final evidence for round 1: tests, smoke, default bench (with CPU baseline + gups), file-to-histogram
# CLI timing, launch list, full capture of the dominant kernel (each ncu pass after the same command
# exited 0 without ncu).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gups > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench.err
# file -> .histo through the C++ driver: multi-threaded reader vs the reference-shaped serial reader
python - <<'PY' > gpurun_out/cli_e2e.txt 2>&1
import os, subprocess, time
from oracle import oracle as o
fq = '/tmp/c2_4m.fastq'
t = time.time(); o.synth_fastq(fq, seed=2, genome_len=20_000_000, read_len=150, sub_rate=0.01, n_rate=0.001, first=0, n=4_000_000)
print('synth 4M reads: %.1f s, %.2f GB' % (time.time() - t, os.path.getsize(fq) / 1e9))
open(fq, 'rb').read()  # page cache
cli = 'sharkmer_b200/host/sharkmer_b200_cli'
os.makedirs('/tmp/cli_out', exist_ok=True)
for name, extra in (('parallel', []), ('parallel', []), ('serial', ['--serial']), ('threads8', ['-t', '8'])):
    t = time.time()
    r = subprocess.run([cli, '-k', '21', '--chunks', '10', '--capacity-hint', '150000000', '-s', name, '-o', '/tmp/cli_out/', *extra, fq], capture_output=True, text=True)
    dt = time.time() - t
    print('%-9s rc=%d wall %.2f s = %.2f M reads/s = %.2f G k-mers/s | %s' % (name, r.returncode, dt, 4.0 / dt, 4e6 * 130 / dt / 1e9, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''))
a = open('/tmp/cli_out/parallel.histo', 'rb').read(); b = open('/tmp/cli_out/serial.histo', 'rb').read()
print('histo identical across readers:', a == b, len(a))
PY
cat gpurun_out/cli_e2e.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"insert_runs|bucket_scatter" -s 30 -c 4 -o gpurun_out/prof_r01c_insert $CMD > gpurun_out/ncu_insert.log 2>&1
echo "insert capture rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','stage_ms','roofline','e2e','cpu_baseline','gups','clocks','gpu_launches')}, indent=1)[:3000])
PY
