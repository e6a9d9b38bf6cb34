#!/bin/bash
# Round 2 call M: pipelined node kernel -- parity (staged == persistent, golden replays), A/B in the bench, smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_staged.py tests/test_permute.py tests/test_engine_parity_gpu.py tests/test_baseline_size.py tests/test_full_size_properties.py -m gpu -x -q ) > gpurun_out/r2m_tests.log 2>&1
tail -n 6 gpurun_out/r2m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -n 2 gpurun_out/r2m_smoke.log
( bash tools/sweep_vlib2.sh 4096 nonodepf main; bash tools/sweep_vlib2.sh 1024 nonodepf main ) > gpurun_out/r2m_sweep.log 2>&1
cat gpurun_out/r2m_sweep.log
