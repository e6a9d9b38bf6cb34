#!/bin/bash
# Round 2 call B: staged kernel chain -- parity vs the persistent kernel, then A/B timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py -m gpu -x -q ) > gpurun_out/r2b_staged_tests.log 2>&1
tail -n 15 gpurun_out/r2b_staged_tests.log
bash tools/sweep_vlib2.sh 1024 main@persist main l256x1 l384x1 l256x3 n256x3 n512x2 > gpurun_out/r2b_sweep.log 2>&1
cat gpurun_out/r2b_sweep.log
bash tools/sweep_vlib2.sh 4096 main > gpurun_out/r2b_sweep4096.log 2>&1
cat gpurun_out/r2b_sweep4096.log
