#!/bin/bash
# Round 2 call V: pollutant pairs in one quality sweep -- parity subset and the default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py tests/test_engine_parity_gpu.py tests/test_continuity.py tests/test_controls.py tests/test_report.py -m gpu -x -q ) > gpurun_out/r2v_tests.log 2>&1
tail -n 5 gpurun_out/r2v_tests.log
( time timeout 1200 python bench.py ) > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
grep "real" gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d[k])
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['phase_ms'])
print('weak', d.get('weak_512_per_gpu')); print('c2', d['c2_single']['kernel_s'], 'c5', d['c5']['kernel_s_max_over_ranks'])
print('cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if 'sample' not in k})
PY
