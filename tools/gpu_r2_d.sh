#!/bin/bash
# Round 2 call D: staged parity tests (full log), per-kernel ncu of one staged routing step (selected kernels)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py -m gpu -q ) > gpurun_out/r2d_staged_tests.log 2>&1
tail -n 60 gpurun_out/r2d_staged_tests.log
timeout 900 ncu --set full --profile-from-start off --clock-control none --import-source on \
    -k regex:'sg_links_pf|sg_nodes|sg_prologue|sg_stream' -c 14 -f -o gpurun_out/r2d_step \
    python tools/profile_launch.py --members 1024 --routing-steps 1 --lib vlib/libswmm_b200_pf256x3.so --out gpurun_out/r2d_step.json > gpurun_out/r2d_ncu.log 2>&1
tail -n 2 gpurun_out/r2d_ncu.log
ncu -i gpurun_out/r2d_step.ncu-rep --page raw --csv > gpurun_out/r2d_step_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_step.ncu-rep --page source --csv --print-source sass -k regex:sg_links_pf > gpurun_out/r2d_links_src.csv 2>/dev/null
ls -la gpurun_out/r2d_step.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2d_step_raw.csv')))
hdr=rows[0]
want=["Kernel Name","gpu__time_duration.sum","launch__registers_per_thread","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct","sm__warps_active.avg.pct_of_peak_sustained_active","l1tex__t_sector_hit_rate.pct","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__inst_executed.sum"]
idx=[hdr.index(w) for w in want if w in hdr]
print([hdr[i] for i in idx])
for r in rows[2:]:
    print([r[i][:28] for i in idx])
PY
rm -f gpurun_out/r2d_step.ncu-rep
