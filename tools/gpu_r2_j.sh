#!/bin/bash
# Round 2 call J: member re-enumeration -- parity on the GPU, A/B in the bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_permute.py tests/test_staged.py -m gpu -x -q ) > gpurun_out/r2j_tests.log 2>&1
tail -n 6 gpurun_out/r2j_tests.log
for RE in 0 100 50; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak --reorder-every $RE 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reorder-every $RE', '%.3e'%d['value'], d['config']['member_reorder'], {k:round(v) for k,v in d['roofline']['phase_ms'].items() if v})"
done > gpurun_out/r2j_ab.log 2>&1
cat gpurun_out/r2j_ab.log
