#!/bin/bash
# Round-2 ncu evidence of the bench's timed launch sequence (4 096 members, 10 routing steps = ~400 kernels):
#   1. plain run of the bracketed launch (event time, conduit-updates, phase timers)
#   2. launch list + DRAM bytes + instructions of EVERY kernel of the sequence (few passes per kernel)
#   3. `--set full` of the dominant kernel (sg_links_pf<0>, first all-members trial) with SASS source page
# Output: gpurun_out/r2_* ; tools/ncu_sequence_summary.py turns 1 + 2 into profiles/ncu_traffic_r02.json.
mkdir -p gpurun_out
M=${1:-4096}
python tools/profile_launch.py --members $M --out gpurun_out/r2_launch_plain.json > gpurun_out/r2_launch_plain.log 2>&1
cat gpurun_out/r2_launch_plain.json; echo
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread \
    --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_sequence.csv \
    python tools/profile_launch.py --members $M --out gpurun_out/r2_launch_ncu.json > gpurun_out/r2_sequence.log 2>&1
tail -n 2 gpurun_out/r2_sequence.log
timeout 900 ncu --set full --profile-from-start off --clock-control none --import-source on \
    -k regex:sg_links_pf -c 1 -f -o gpurun_out/r2_links_full \
    python tools/profile_launch.py --members $M --routing-steps 1 --out gpurun_out/r2_links_full_launch.json > gpurun_out/r2_links_full.log 2>&1
tail -n 2 gpurun_out/r2_links_full.log
ncu -i gpurun_out/r2_links_full.ncu-rep --page raw --csv > gpurun_out/r2_links_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_links_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_links_full_src.csv 2>/dev/null
gzip -f gpurun_out/r2_links_full_src.csv
rm -f gpurun_out/r2_links_full.ncu-rep
python tools/ncu_sequence_summary.py gpurun_out/r2_sequence.csv gpurun_out/r2_launch_plain.json gpurun_out/r2_traffic.json
