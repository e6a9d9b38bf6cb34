#!/bin/bash
mkdir -p gpurun_out
bash tools/sweep_vlib.sh main stream stream_tile2 > gpurun_out/r2_sweep4.log 2>&1
cat gpurun_out/r2_sweep4.log
