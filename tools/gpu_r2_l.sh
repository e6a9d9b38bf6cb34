#!/bin/bash
# Round 2 call L (4 GPUs): config 5 striped over 4 GPUs with the bit-for-bit check, bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    tools/c5_partitioned.py --check --sim-s 3600 ) > gpurun_out/r2l_c5_n4.json 2> gpurun_out/r2l_c5_n4.err
tail -n 3 gpurun_out/r2l_c5_n4.err; cut -c1-1500 gpurun_out/r2l_c5_n4.json | tail -n 2
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 10 --warmup 3 --no-packed ) > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err
grep "bench \|real\|rror" gpurun_out/r2l_bench_n4.err | tail -n 12
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2l_bench_n4.json').read().strip().splitlines()[-1])
    for k in ('value','n_gpus','ms_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
    print('weak', d.get('weak_512_per_gpu')); print('c5', d.get('c5'))
except Exception as e: print('failed', e)
PY
