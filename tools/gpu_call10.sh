#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main tile2 s2 tile2s2 > gpurun_out/r2_sweep10.log 2>&1
cat gpurun_out/r2_sweep10.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
tail -n 5 gpurun_out/r2_bench10.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench10.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','picard_iterations_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['phase_ms'])
    print('weak', d.get('weak_512_per_gpu')); print('c2', d.get('c2_single')); print('c5', d.get('c5')); print('cpu', d.get('cpu_baseline'))
except Exception as e: print('bench failed', e)
PY
