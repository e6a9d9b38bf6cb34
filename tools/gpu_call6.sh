#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main nopersist stream > gpurun_out/r2_sweep6.log 2>&1
cat gpurun_out/r2_sweep6.log
timeout 900 python -m pytest tests/test_engine_parity_gpu.py -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1
tail -n 5 gpurun_out/r2_pytest6.log
