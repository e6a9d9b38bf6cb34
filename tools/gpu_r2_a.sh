#!/bin/bash
# Round 2, re-entry call A: GPU tests, default bench, reference arm, launch list + full capture of the C4 launch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/r2a_env.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_gputests.log 2>&1
tail -n 5 gpurun_out/r2a_gputests.log
( time timeout 900 python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
grep "bench \|real" gpurun_out/r2a_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
grep real gpurun_out/r2a_ref.err
args="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2a_launches.csv python bench.py $args > gpurun_out/r2a_ncu_list.log 2>&1
tail -n 4 gpurun_out/r2a_launches.csv
python tools/profile_launch.py --members 4096 --out gpurun_out/r2a_launch_plain.json > gpurun_out/r2a_launch_plain.log 2>&1
timeout 1200 ncu --set full --profile-from-start off --clock-control none --import-source on -f -o gpurun_out/r2a_prof \
    python tools/profile_launch.py --members 4096 --out gpurun_out/r2a_launch_ncu.json > gpurun_out/r2a_ncu_full.log 2>&1
tail -n 3 gpurun_out/r2a_ncu_full.log
ncu -i gpurun_out/r2a_prof.ncu-rep --page raw --csv > gpurun_out/r2a_prof_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','picard_iterations_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['phase_ms'])
    print('weak', d.get('weak_512_per_gpu')); print('c2', d.get('c2_single')); print('c5', d.get('c5')); print('cpu', d.get('cpu_baseline'))
    r=json.loads(open('gpurun_out/r2a_ref.json').read().strip().splitlines()[-1])
    print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'])
except Exception as e: print('failed', e)
PY
