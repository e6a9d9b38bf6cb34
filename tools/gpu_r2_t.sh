#!/bin/bash
# Round 2 call T: e2e (host buffers every step) against the number of member blocks in flight
mkdir -p gpurun_out
for B in 8 16; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-c2-single --no-weak --e2e-blocks $B --e2e-steps 12 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('blocks $B', 'e2e %.3e'%d['e2e']['value'], 'device %.3e'%d['value'])"
done | tee gpurun_out/r2t_e2e_blocks.log
