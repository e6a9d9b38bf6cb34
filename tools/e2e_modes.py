#!/usr/bin/env python
"""Compares the host-buffer step modes on the bench workload: swb_step_host on the whole ensemble
(device-chosen dt / host-fed dt) and swb_step_host_batch over 2, 4, 8 member blocks."""
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = types.SimpleNamespace(grid=100, hours=6.0, surcharge="SLOT", members=512, members_total=512,
                             spinup=float(os.environ.get("SPINUP", "6000")), e2e_steps=10, e2e_blocks=1)
s, case, spec = bench.make_ensemble(args, 0, 0)
bench.spin_up(s, args.spinup)
n_true = int(case.net.true_conduit_mask().sum())
net = case.net
M, nP = s.M, net.n_pollut
lat = s.host_array((M, net.n_nodes)); lat[:] = s.get_field("SWB_NODE_NEW_LATFLOW")
conc = np.zeros((net.n_nodes, nP)); conc[case.inflows["node"]] = case.inflows["concen"].reshape(-1, nP)
load = s.host_array((M, net.n_nodes, nP)); load[:] = np.maximum(lat, 0.0)[:, :, None] * conc[None]
depth = s.host_array((M, net.n_nodes)); flow = s.host_array((M, net.n_links))
next_dt = s.host_array((M,)); iters = s.host_array((M,), dtype=np.int32)
for mode in ("device_dt", "host_dt"):
    dt = s.host_array((M,)); dt[:] = [x.next_dt for x in s.stats(0, M)]
    kw = dict(qual_load=load, node_depth=depth, link_flow=flow, next_dt=next_dt, iters=iters)
    if mode == "host_dt":
        kw["dt"] = dt
    s.step_host(lat, **kw)
    cu0 = s.conduit_updates(); t0 = time.perf_counter()
    for _ in range(10):
        s.step_host(lat, **kw)
        dt[:] = next_dt
    sec = time.perf_counter() - t0
    cu = s.conduit_updates() - cu0
    print(mode, "cu/step %.3e" % (cu / 10), "ms/step %.2f" % (sec * 100), "cu/s %.3e" % (cu / sec), "kernel ms", s.last_kernel_ms())
for blocks in (2, 4, 8):
    args.e2e_blocks = blocks
    r = bench.measure_e2e(s, case, args, n_true)
    print("batch", blocks, "cu/step %.3e" % (r["cu"] / r["steps"]), "ms/step %.2f" % (r["seconds"] * 1000 / r["steps"]),
          "cu/s %.3e" % (r["cu"] / r["seconds"]))
# kernel time of one member block stepped alone
nb = 128
subs = [s.clone_members(b * nb, nb) for b in range(4)]
dt = s.host_array((M,)); dt[:] = [x.next_dt for x in s.stats(0, M)]
for rep in range(3):
    ks = []
    t0 = time.perf_counter()
    for b, x in enumerate(subs):
        sl = slice(b * nb, (b + 1) * nb)
        x.step_host(lat[sl], dt=dt[sl], qual_load=load[sl], node_depth=depth[sl], link_flow=flow[sl],
                    next_dt=next_dt[sl], iters=iters[sl])
        ks.append(round(x.last_kernel_ms(), 3))
    print("sequential blocks: kernel ms", ks, "total ms %.2f" % ((time.perf_counter() - t0) * 1000))
    dt[:] = next_dt
