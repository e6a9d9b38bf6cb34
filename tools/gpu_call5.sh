#!/bin/bash
mkdir -p gpurun_out
python tools/profile_phase.py links > gpurun_out/r2_phase_links.txt 2>&1
python tools/profile_phase.py nodes > gpurun_out/r2_phase_nodes.txt 2>&1
cat gpurun_out/r2_phase_links.txt gpurun_out/r2_phase_nodes.txt
for ph in links nodes; do
  ncu --set full --profile-from-start off --clock-control none --import-source on -f -o gpurun_out/r2_$ph python tools/profile_phase.py $ph > gpurun_out/r2_ncu_$ph.log 2>&1
  tail -n 3 gpurun_out/r2_ncu_$ph.log
done
ls -la gpurun_out/*.ncu-rep
