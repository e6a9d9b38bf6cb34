#!/bin/bash
# Round 2 call S: timed region of the default bench (20 x 10 routing steps), final tree
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], d['roofline']['frac'], d['config']['member_reorder'][-60:], {k:round(v) for k,v in d['roofline']['phase_ms'].items() if v})"
done
