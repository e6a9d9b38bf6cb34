#!/bin/bash
# usage: bash tools/build_variant.sh <name> [-DSWB_... flags]   -> vlib/libswmm_b200_<name>.so (+ scratch/ptxas_<name>.log)
name=$1; shift
C=stormwater-management-model_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" -I $C -I include $C/swb_api.cu -o vlib/libswmm_b200_$name.so > scratch/ptxas_$name.log 2>&1
grep -A2 "swb_route_kernel" scratch/ptxas_$name.log | grep -i "registers\|spill" | head -4
