#!/bin/bash
# Round 2 call W: launch list + DRAM bytes of every kernel of one bench launch sequence, final tree
mkdir -p gpurun_out
python tools/profile_launch.py --members 4096 --out gpurun_out/r2_launch_plain.json > gpurun_out/r2_launch_plain.log 2>&1
cat gpurun_out/r2_launch_plain.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread \
    --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_sequence.csv \
    python tools/profile_launch.py --members 4096 --out gpurun_out/r2_launch_ncu.json > gpurun_out/r2_sequence.log 2>&1
tail -n 1 gpurun_out/r2_sequence.log | cut -c1-200
python tools/ncu_sequence_summary.py gpurun_out/r2_sequence.csv gpurun_out/r2_launch_plain.json gpurun_out/r2_traffic.json | cut -c1-260
