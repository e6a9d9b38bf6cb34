#!/bin/bash
# Round 2 call K: final-tree evidence -- whole GPU suite, default bench (both arms), ncu launch sequence + full capture
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2k_gputests.log 2>&1
tail -n 6 gpurun_out/r2k_gputests.log
( time timeout 1200 python bench.py ) > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
grep "bench \|real" gpurun_out/r2k_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2k_reference.json 2> gpurun_out/r2k_reference.err
grep "real" gpurun_out/r2k_reference.err; cut -c1-400 gpurun_out/r2k_reference.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','picard_iterations_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['phase_ms'])
    print('dominant', {k: v for k, v in d['roofline']['dominant_kernel'].items() if k != 'ncu'})
    print('reorder', d['config']['member_reorder'])
    print('weak', d.get('weak_512_per_gpu')); print('c2', d.get('c2_single')); print('c5', d.get('c5'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e: print('failed', e)
PY
bash tools/ncu_capture_r2.sh 4096 > gpurun_out/r2k_ncu_capture.log 2>&1
tail -n 16 gpurun_out/r2k_ncu_capture.log | cut -c1-300
