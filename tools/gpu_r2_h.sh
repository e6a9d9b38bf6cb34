#!/bin/bash
# Round 2 call H: whole GPU suite, occupancy A/B of the link kernel with staged statics, bench, drop-in timing
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_gputests.log 2>&1
tail -n 8 gpurun_out/r2h_gputests.log
( bash tools/sweep_vlib2.sh 4096 main pf256x4 pf256x2 ) > gpurun_out/r2h_sweep.log 2>&1
cat gpurun_out/r2h_sweep.log
( bash tools/dropin_timing.sh 100 2; bash tools/dropin_timing.sh 30 2 ) 2>&1 | grep -v "rpt:" > gpurun_out/r2h_dropin.log
cat gpurun_out/r2h_dropin.log
( time timeout 900 python bench.py --no-packed ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
grep "bench \|real" gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','picard_iterations_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['phase_ms'])
    print('weak', d.get('weak_512_per_gpu')); print('c2', d.get('c2_single')); print('c5', d.get('c5'))
except Exception as e: print('failed', e)
PY
