#!/bin/bash
# Round 2 call C: cp.async-pipelined link kernel -- parity, A/B timing, per-kernel ncu of one staged routing step
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py -m gpu -x -q ) > gpurun_out/r2c_staged_tests.log 2>&1
tail -n 6 gpurun_out/r2c_staged_tests.log
bash tools/sweep_vlib2.sh 1024 main pf256x3 pf128x6 pf256x4 nopf256x3 > gpurun_out/r2c_sweep.log 2>&1
cat gpurun_out/r2c_sweep.log
timeout 900 ncu --set full --profile-from-start off --clock-control none --import-source on -f -o gpurun_out/r2c_step \
    python tools/profile_launch.py --members 1024 --routing-steps 1 --out gpurun_out/r2c_step.json > gpurun_out/r2c_ncu.log 2>&1
tail -n 2 gpurun_out/r2c_ncu.log
ncu -i gpurun_out/r2c_step.ncu-rep --page raw --csv > gpurun_out/r2c_step_raw.csv 2>/dev/null
ls -la gpurun_out/r2c_step.ncu-rep
