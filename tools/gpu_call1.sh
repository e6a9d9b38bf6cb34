#!/bin/bash
# GPU call 1 of round 2: conditional-graph probe, launch-bounds sweep with phase timers, C4 at full width on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/r2_call1_env.txt
./tools/micro/graphwhile > gpurun_out/r2_graphwhile.txt 2>&1
bash tools/sweep_vlib.sh main b256x2 b512x1 b256x3 b384x2 b1024x1 > gpurun_out/r2_sweep1.log 2>&1
python bench.py --members 4096 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_m4096.json 2> gpurun_out/r2_m4096.err
python bench.py --members 2048 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_m2048.json 2> gpurun_out/r2_m2048.err
cat gpurun_out/r2_graphwhile.txt gpurun_out/r2_sweep1.log
python -c "
import json
for f in ('gpurun_out/r2_m4096.json','gpurun_out/r2_m2048.json'):
    try:
        d=json.load(open(f)); print(f, '%.3e'%d['value'], d['ms_per_step'], d['roofline']['phase_ms'])
    except Exception as e: print(f, 'failed', e)
"
