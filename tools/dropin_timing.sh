#!/bin/bash
# Wall time of the unmodified CLI with and without the CUDA seam preloaded (single model).
# usage: tools/dropin_timing.sh <grid n> <hours>
set -e
cd "$(dirname "$0")/.."
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
import swmm_b200
from swmm_b200 import scenarios
open("$D/m.inp", "w").write(scenarios.c2_grid_inp(scenarios.GridSpec(nx=$1, ny=$1, hours=$2, threads=$(nproc))))
PY
export OMP_PROC_BIND=true OMP_WAIT_POLICY=active
dur() { python -c "print('%.2f' % ($2 - $1))"; }
s=$(date +%s.%N); oracle/_ref/runswmm $D/m.inp $D/ref.rpt $D/ref.out > /dev/null; e=$(date +%s.%N)
echo "reference CLI ($(nproc) threads): $(dur $s $e) s"
s=$(date +%s.%N); LD_PRELOAD=$PWD/stormwater-management-model_b200/seam/libswmm5_b200_seam.so oracle/_ref/runswmm $D/m.inp $D/gpu.rpt $D/gpu.out > /dev/null; e=$(date +%s.%N)
echo "CLI + B200 seam: $(dur $s $e) s"
grep -A3 "Flow Routing Continuity" $D/ref.rpt | head -1 > /dev/null
grep "Continuity Error" $D/ref.rpt $D/gpu.rpt
