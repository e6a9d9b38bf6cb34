#!/bin/bash
# Round profile of the default bench step (run on the GPU box through gpurun):
#   1. the plain command must exit 0 first; 2. launch list (gpu__time_duration per launch);
#   3. one --set full capture of a timed swb_route_kernel launch, exported as raw CSV.
# usage: bash tools/ncu_capture.sh <tag>        -> gpurun_out/{bench,launches,prof}_<tag>.*
tag=${1:-r01}
args="--steps 2 --warmup 1 --spinup 6000 --no-cpu-baseline --e2e-steps 2"
mkdir -p gpurun_out
python bench.py $args > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $args > gpurun_out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swb_route -s 28 -c 1 \
    -o gpurun_out/prof_$tag -f python bench.py $args > gpurun_out/ncu_full_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_$tag.ncu-rep
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/prof_${tag}_raw.csv")))
hdr, val = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "lts__t_sectors_srcunit_tex_aperture_device_op_read_lookup_hit.sum"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(w, rows[1][i], val[i])
PY
