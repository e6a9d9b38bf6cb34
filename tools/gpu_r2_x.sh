#!/bin/bash
# Round 2 call X: default bench of the final tree with the final ncu traffic file
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
grep "real" gpurun_out/r2x_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
