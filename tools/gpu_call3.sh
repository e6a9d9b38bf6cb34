#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main > gpurun_out/r2_sweep3.log 2>&1
cat gpurun_out/r2_sweep3.log
timeout 1500 python -m pytest tests/test_engine_parity_gpu.py tests/test_baseline_size.py -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1
tail -n 8 gpurun_out/r2_pytest3.log
