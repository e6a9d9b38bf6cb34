#!/bin/bash
# Round 2 call F: device-side controls / general inflows on the GPU, drop-in wall times, per-line ncu of the
# link and node kernels of one staged routing step
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_controls.py -m gpu -x -q ) > gpurun_out/r2f_controls.log 2>&1
tail -n 12 gpurun_out/r2f_controls.log
( bash tools/dropin_timing.sh 100 2; bash tools/dropin_timing.sh 30 2 ) > gpurun_out/r2f_dropin.log 2>&1
cat gpurun_out/r2f_dropin.log
for K in sg_links_pf sg_nodes; do
  timeout 600 ncu --set full --profile-from-start off --clock-control none --import-source on \
      -k regex:$K -c 1 -f -o gpurun_out/r2f_$K \
      python tools/profile_launch.py --members 1024 --routing-steps 1 --out gpurun_out/r2f_$K.json > gpurun_out/r2f_ncu_$K.log 2>&1
  tail -n 2 gpurun_out/r2f_ncu_$K.log
  ncu -i gpurun_out/r2f_$K.ncu-rep --page raw --csv > gpurun_out/r2f_${K}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2f_$K.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_${K}_src.csv 2>/dev/null
  gzip -f gpurun_out/r2f_${K}_src.csv
  rm -f gpurun_out/r2f_$K.ncu-rep
done
ls -la gpurun_out | grep r2f
