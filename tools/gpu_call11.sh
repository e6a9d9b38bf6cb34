#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
grep "bench \|real" gpurun_out/r2_bench11.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r2_ref11.json 2> gpurun_out/r2_ref11.err
grep real gpurun_out/r2_ref11.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench11.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step'): print(k, d[k])
    print('c5', d['c5']['kernel_s_max_over_ranks'], d['c5']['conduit_updates_per_s'])
    print('cpu', d.get('cpu_baseline'))
    r=json.loads(open('gpurun_out/r2_ref11.json').read().strip().splitlines()[-1])
    print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'])
except Exception as e: print('failed', e)
PY
