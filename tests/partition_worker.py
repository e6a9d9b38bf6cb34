"""One rank of the partitioned-network parity test (launched by torchrun from test_partition.py):
runs its stripe of a golden grid, gathers the owned results on rank 0 and compares them there with
the single-domain run of the same library.  SWB_LIB selects the host emulation (CPU suite); with
--cuda the CUDA library runs, one device per rank when the box has enough of them."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import parity_common as pc
from swmm_b200 import partition, solver
from test_partition import FIELDS, golden_setup


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--case", default="c2_grid12_slot")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--cuda", action="store_true")
    a = ap.parse_args()
    dist.init_process_group(a.backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = None if a.cuda else os.environ["SWB_LIB"]
    device = 0
    if a.cuda:
        ndev = solver.load_library().swb_device_count()
        device = rank % max(ndev, 1)
    net, state, inflows, t_end = golden_setup(a.case)
    n = int(round((net.n_nodes - 1) ** 0.5))
    parts = partition.split_network(net, partition.stripes(n, n, world, extra_nodes=1), world)
    ps = partition.PartitionedSolver(parts[rank], device=device, lib_path=lib, timeout_s=20.0)
    handles = [None] * world
    dist.all_gather_object(handles, ps.export_handle())
    ps.connect(handles)
    ps.load_state(partition.split_state(ps.part, state, net.n_pollut))
    ps.set_inflows(**partition.split_inflows(ps.part, inflows, net.n_pollut))
    dist.barrier()
    done = 0
    for chunk in (1, 2, 7, a.steps):
        ps.run_steps(chunk, t_end)
        done += chunk
    pieces = {f: ps.owned_field(f) for f in FIELDS}
    st = ps.stats()[0]
    mine = dict(pieces=pieces, sim_time=st.sim_time, iterations=st.iterations, next_dt=st.next_dt,
                exchanges=ps.exchanges())
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    ok = True
    if rank == 0:
        single = solver.Solver(net, 1, device=device, lib_path=lib)
        single.load_state(state)
        single.set_inflows(**inflows)
        single.run_steps(done, t_end)
        s0 = single.stats()[0]
        for p in allp:
            ok = ok and p["sim_time"] == s0.sim_time and p["iterations"] == s0.iterations and p["next_dt"] == s0.next_dt
        for f in FIELDS:
            w = net.n_pollut if f.endswith("_QUAL") else 1
            n_items = net.n_nodes if f.startswith("SWB_NODE") else net.n_links
            got = partition.assemble([p["pieces"][f] for p in allp], n_items, w)
            ref = single.get_field(f)[0]
            if not np.array_equal(got, ref):
                ok = False
                print("MISMATCH", f, float(np.max(np.abs(got - ref))))
        print(f"steps {done} iterations {s0.iterations} exchanges {allp[0]['exchanges']} sim_time {s0.sim_time}")
        print("partition parity ok" if ok else "partition parity FAILED")
        single.close()
    ps.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
