#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main nopersist > gpurun_out/r2_sweep7.log 2>&1
cat gpurun_out/r2_sweep7.log
