#!/bin/bash
# Round 2 call I: outfall depths in the link phase, slim seam -- parity subset, member trial statistics,
# drop-in timing, C5 on one GPU, ncu evidence of the bench launch sequence
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_staged.py tests/test_engine_parity_gpu.py tests/test_partition.py tests/test_seam_dropin.py -m gpu -x -q ) > gpurun_out/r2i_tests.log 2>&1
tail -n 6 gpurun_out/r2i_tests.log
timeout 600 python tools/member_trials.py --members 1024 --out gpurun_out/r2i_member_trials.npz > gpurun_out/r2i_member_trials.json 2> gpurun_out/r2i_member_trials.err
cat gpurun_out/r2i_member_trials.json
( bash tools/dropin_timing.sh 100 2; bash tools/dropin_timing.sh 30 2 ) 2>&1 | grep -v "rpt:" > gpurun_out/r2i_dropin.log
cat gpurun_out/r2i_dropin.log
timeout 600 python tools/c5_single_gpu.py > gpurun_out/r2i_c5_single.log 2>&1
tail -n 4 gpurun_out/r2i_c5_single.log
bash tools/ncu_capture_r2.sh 4096 > gpurun_out/r2i_ncu_capture.log 2>&1
tail -n 25 gpurun_out/r2i_ncu_capture.log
ls -la gpurun_out | grep "r2_"
