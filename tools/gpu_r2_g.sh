#!/bin/bash
# Round 2 call G: static rows staged in shared memory (link kernel) -- parity, A/B against the previous build,
# member order A/B, device-side controls on the GPU
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py -m gpu -x -q ) > gpurun_out/r2g_staged.log 2>&1
tail -n 6 gpurun_out/r2g_staged.log
( time timeout 1200 python -m pytest tests/test_controls.py -m gpu -x -q ) > gpurun_out/r2g_controls.log 2>&1
tail -n 8 gpurun_out/r2g_controls.log
( bash tools/sweep_vlib2.sh 1024 prev main; bash tools/sweep_vlib2.sh 4096 prev main; EXTRA="--member-order scale" bash tools/sweep_vlib2.sh 4096 main ) > gpurun_out/r2g_sweep.log 2>&1
cat gpurun_out/r2g_sweep.log
( bash tools/dropin_timing.sh 100 2; bash tools/dropin_timing.sh 30 2 ) > gpurun_out/r2g_dropin.log 2>&1
cat gpurun_out/r2g_dropin.log
