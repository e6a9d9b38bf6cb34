#!/bin/bash
# Round 2 call N (4 GPUs): config 5 over 4 GPUs (link phase starts at the outfall's conduit), bench.py under torchrun
# with interleaved members, two solvers on two devices in one process
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_permute.py tests/test_partition.py -m gpu -x -q ) > gpurun_out/r2n_tests.log 2>&1
tail -n 4 gpurun_out/r2n_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    tools/c5_partitioned.py --check --sim-s 3600 ) > gpurun_out/r2n_c5_n4.json 2> gpurun_out/r2n_c5_n4.err
tail -n 3 gpurun_out/r2n_c5_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2n_c5_n4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('kernel_s_max_over_ranks','conduit_updates_per_s','single_gpu_kernel_s','identical_to_single_gpu')})
for r in d['phase_ms_per_rank']: print({k: round(v) for k, v in r.items() if v})"
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --no-packed ) > gpurun_out/r2n_bench_n4.json 2> gpurun_out/r2n_bench_n4.err
grep "real\|rror" gpurun_out/r2n_bench_n4.err | tail -n 5
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2n_bench_n4.json').read().strip().splitlines()[-1])
    for k in ('value','n_gpus','ms_per_step','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
    print('weak', d.get('weak_512_per_gpu')); print('c5', {k: v for k, v in d.get('c5', {}).items() if k != 'phase_ms_per_rank'})
except Exception as e: print('failed', e)
PY
