#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh p1 main > gpurun_out/r2_sweep9.log 2>&1
cat gpurun_out/r2_sweep9.log
timeout 900 python -m pytest tests/test_engine_parity_gpu.py tests/test_baseline_size.py -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1
tail -n 5 gpurun_out/r2_pytest9.log
