#!/bin/bash
# Round 2 call R: interface-file inflows and the control tests on the GPU (the prologue changed)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_controls.py tests/test_staged.py tests/test_continuity.py -m gpu -x -q ) > gpurun_out/r2r_tests.log 2>&1
tail -n 6 gpurun_out/r2r_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], d['roofline']['frac'])"
