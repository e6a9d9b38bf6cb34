"""One rank of tests/test_c5_golden.py::test_cuda_config5_two_gpus_vs_reference_fixture (torchrun): config 5
striped over the ranks, replayed against the reference fixture at its snapshot steps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402
from swmm_b200 import network, partition, scenarios  # noqa: E402
from test_c5_golden import FIELDS, TOL, _compare, _fixture  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    g = _fixture()
    nx, ny = int(g["nx"]), int(g["ny"])
    spec = scenarios.GridSpec(nx=nx, ny=ny, hours=float(g["sim_min"]) / 60.0, pollutants=False, surcharge="SLOT")
    case = network.build_grid(spec)
    net = case.net
    parts = partition.split_network(net, partition.stripes(ny, nx, world, extra_nodes=1), world)
    ps = partition.PartitionedSolver(parts[rank], device=local, timeout_s=60.0)
    handles = [None] * world
    dist.all_gather_object(handles, ps.export_handle())
    ps.connect(handles)
    ps.load_state(partition.split_state(ps.part, case.state0, net.n_pollut))
    ps.set_inflows(**partition.split_inflows(ps.part, case.inflows, net.n_pollut))
    dist.barrier()
    times, iters = g["series_time"], g["series_iters"]
    worst, done, ok = {}, 0, True
    for k, step in enumerate(g["snap_steps"]):
        ps.run_steps(int(step) - done, case.t_end)
        done = int(step)
        st = ps.stats()[0]
        ok = ok and abs(st.sim_time - times[done - 1]) < 1e-9 and st.iterations == int(np.sum(iters[:done]))
        pieces = {f: ps.owned_field(f) for f in FIELDS}
        allp = [None] * world
        dist.all_gather_object(allp, pieces)
        if rank == 0:
            full = {}
            for f in FIELDS:
                n_items = net.n_nodes if f.startswith("SWB_NODE") else net.n_links
                full[f] = partition.assemble([p[f] for p in allp], n_items, 1)
            _compare(lambda f: full[f], g, k, worst)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(worst)
        good = bool(flag.item()) and done == len(times) and max(worst.values()) <= TOL
        print("C5 GOLDEN OK" if good else "C5 GOLDEN FAILED")
    ps.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
