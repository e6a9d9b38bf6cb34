#!/bin/bash
# Round 2 call U: default bench of the final tree (e2e with 16 member blocks)
mkdir -p gpurun_out
( time timeout 1200 python bench.py ) > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
grep "real" gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d[k])
print('e2e', d['e2e']['value'], d['e2e']['member_blocks_per_gpu'], 'roofline', d['roofline']['frac'])
print('cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if 'sample' not in k})
PY
