#!/bin/bash
# Round 2 call O (4 GPUs): config 5 over 4 GPUs, outfall conduits on threads of their own
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_partition.py -m gpu -x -q ) > gpurun_out/r2o_tests.log 2>&1
tail -n 3 gpurun_out/r2o_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    tools/c5_partitioned.py --check --sim-s 3600 ) > gpurun_out/r2o_c5_n4.json 2> gpurun_out/r2o_c5_n4.err
tail -n 3 gpurun_out/r2o_c5_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2o_c5_n4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('kernel_s_max_over_ranks','conduit_updates_per_s','single_gpu_kernel_s','identical_to_single_gpu')})
for r in d['phase_ms_per_rank']: print({k: round(v) for k, v in r.items() if v})"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tools/c5_partitioned.py --check --sim-s 3600 ) > gpurun_out/r2o_c5_n2.json 2> gpurun_out/r2o_c5_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2o_c5_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('kernel_s_max_over_ranks','conduit_updates_per_s','single_gpu_kernel_s','identical_to_single_gpu')})"
