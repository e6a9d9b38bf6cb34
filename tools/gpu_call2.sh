#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main tile1 tile2 lookni > gpurun_out/r2_sweep2.log 2>&1
cat gpurun_out/r2_sweep2.log
timeout 1200 python -m pytest tests/test_baseline_size.py tests/test_conditioning.py -m gpu -x -q -s > gpurun_out/r2_pytest_size.log 2>&1
tail -n 30 gpurun_out/r2_pytest_size.log
