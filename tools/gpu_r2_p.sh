#!/bin/bash
# Round 2 call P: final tree -- whole GPU suite, smoke, default bench of both arms
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2p_gputests.log 2>&1
tail -n 6 gpurun_out/r2p_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -n 1 gpurun_out/r2p_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
grep "real" gpurun_out/r2p_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2p_reference.json 2> gpurun_out/r2p_reference.err
grep "real" gpurun_out/r2p_reference.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','picard_iterations_per_step','gpu_launches'): print(k, d[k])
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['phase_ms'])
print('dominant', {k: v for k, v in d['roofline']['dominant_kernel'].items() if k != 'ncu'})
print('reorder', d['config']['member_reorder'])
print('weak', d.get('weak_512_per_gpu')); print('c2', d.get('c2_single')); print('c5', {k: v for k, v in d.get('c5', {}).items() if k != 'workload'})
print('cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if 'sample' not in k})
r=json.loads(open('gpurun_out/r2p_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'])
PY
