#!/bin/bash
# Round 2 call Q: the new robustness tests on the GPU
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_permute.py -m gpu -x -q ) > gpurun_out/r2q_tests.log 2>&1
tail -n 8 gpurun_out/r2q_tests.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
