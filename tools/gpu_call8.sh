#!/bin/bash
mkdir -p gpurun_out
for m in 32 64 128 256; do
python bench.py --members $m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('M=$m', '%.3e'%d['value'], {k:round(v,1) for k,v in d['roofline']['phase_ms'].items()})"
done > gpurun_out/r2_sweep8.log 2>&1
cat gpurun_out/r2_sweep8.log
