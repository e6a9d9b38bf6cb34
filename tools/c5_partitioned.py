#!/usr/bin/env python
"""BASELINE config 5: the 1000 x 500 looped grid (998 501 conduits, SLOT) as ONE model cut into
stripes over N GPUs, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/c5_partitioned.py [--nx 1000 --ny 500 --sim-s 600 --check]

The border depths travel inside the persistent kernel (peer windows over NVLink, csrc/swb_engine.h:
halo_exchange); torch.distributed carries only the 64-byte window handles (gloo) and the final
result gather (NCCL all_gather of the owned depths).  --check makes rank 0 run the same model
unpartitioned on its own GPU afterwards and compare every depth and flow bit for bit.
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swmm_b200  # noqa: F401,E402
from swmm_b200 import network, partition, scenarios, solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--sim-s", type=float, default=600.0)
    ap.add_argument("--pollutants", action="store_true")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    t0 = time.perf_counter()
    spec = scenarios.GridSpec(nx=a.nx, ny=a.ny, hours=1.0, pollutants=a.pollutants, surcharge="SLOT")
    case = network.build_grid(spec)
    net = case.net
    nP = net.n_pollut
    parts = partition.split_network(net, partition.stripes(a.ny, a.nx, world, extra_nodes=1), world)
    part = parts[rank]
    del parts
    build_s = time.perf_counter() - t0
    ps = partition.PartitionedSolver(part, device=local, timeout_s=20.0)
    handles = [None] * world
    dist.all_gather_object(handles, ps.export_handle())
    ps.connect(handles)
    ps.load_state(partition.split_state(part, case.state0, nP))
    ps.set_inflows(**partition.split_inflows(part, case.inflows, nP))
    dist.barrier()
    ps.run_steps(5, a.sim_s)                  # warm-up launch (also pages the peer mappings in)
    it0 = ps.stats()[0].iterations
    ex0 = ps.exchanges()
    ps.phase_times()
    dist.barrier()
    w0 = time.perf_counter()
    ps.run_steps(10_000_000, a.sim_s)         # ONE launch to the end of the simulation
    wall = time.perf_counter() - w0
    ms = torch.tensor([ps.last_kernel_ms()], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = ps.stats()[0]
    phases = ps.phase_times()
    # result gather over NCCL: owned depths, padded to the largest stripe
    n_own = torch.tensor([part.n_owned], device="cuda")
    sizes = [torch.zeros_like(n_own) for _ in range(world)]
    dist.all_gather(sizes, n_own)
    cap = int(max(int(x) for x in sizes))
    mine = torch.zeros(cap, dtype=torch.float64, device="cuda")
    gid, depth = ps.owned_field("SWB_NODE_NEW_DEPTH")
    mine[:part.n_owned] = torch.from_numpy(depth).cuda()
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    gids = [None] * world
    dist.all_gather_object(gids, gid)
    flows = [None] * world
    quals = [None] * world
    if a.check:
        dist.all_gather_object(flows, ps.owned_field("SWB_LINK_NEW_FLOW"))
        if nP:
            dist.all_gather_object(quals, ps.owned_field("SWB_LINK_NEW_QUAL"))
    all_phases = [None] * world
    dist.all_gather_object(all_phases, {k: round(v, 3) for k, v in phases.items()})
    ok = None
    if rank == 0:
        depth_all = np.zeros(net.n_nodes)
        for r in range(world):
            depth_all[gids[r]] = allv[r][:gids[r].size].cpu().numpy()
        n_true = int(net.true_conduit_mask().sum())
        cu = (st.iterations - it0) * n_true
        t = float(ms.item()) / 1000.0
        out = {"workload": f"C5 {a.nx}x{a.ny} looped grid as one model, {n_true} conduits, SLOT, "
                           f"{'2 pollutants' if nP else 'no pollutants'}, striped over {world} GPU(s)",
               "n_gpus": world, "sim_s": st.sim_time, "steps": int(st.steps), "iterations": int(st.iterations),
               "timed_iterations": int(st.iterations - it0), "kernel_s_max_over_ranks": t, "wall_s": wall,
               "conduit_updates_per_s": cu / t, "algorithmic_GBps": 200.0 * cu / t / 1e9,
               "halo_exchanges_timed": ps.exchanges() - ex0,
               "halo_nodes_rank0": int(part.recv_node.size), "phase_ms_per_rank": all_phases,
               "build_s": build_s, "max_depth_ft": float(depth_all.max())}
        if a.check:
            single = solver.Solver(net, 1, device=local)
            single.load_state(case.state0)
            single.set_inflows(**case.inflows)
            single.run_steps(5, a.sim_s)
            single.run_steps(10_000_000, a.sim_s)
            s0 = single.stats()[0]
            ref_d = single.get_field("SWB_NODE_NEW_DEPTH")[0]
            ref_q = single.get_field("SWB_LINK_NEW_FLOW")[0]
            flow_all = partition.assemble(flows, net.n_links)
            ok = bool(np.array_equal(depth_all, ref_d) and np.array_equal(flow_all, ref_q)
                      and s0.iterations == st.iterations and s0.sim_time == st.sim_time and s0.steps == st.steps)
            if nP:      # the border concentrations travel through the same windows (qualrout.c:253-353)
                qual_all = partition.assemble(quals, net.n_links, nP)
                ok = ok and bool(np.array_equal(qual_all, single.get_field("SWB_LINK_NEW_QUAL")[0]))
                out["max_link_concentration"] = float(qual_all.max())
            out["single_gpu_kernel_s"] = single.last_kernel_ms() / 1000.0
            out["single_gpu_conduit_updates_per_s"] = cu / (single.last_kernel_ms() / 1000.0)
            out["identical_to_single_gpu"] = ok
            out["max_abs_depth_diff"] = float(np.max(np.abs(depth_all - ref_d)))
            single.close()
        print(json.dumps(out), flush=True)
    ps.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok in (None, True) else 1)


if __name__ == "__main__":
    main()
