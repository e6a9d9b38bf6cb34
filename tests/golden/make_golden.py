#!/usr/bin/env python
"""Writes tests/golden/<case>.npz from the UNMODIFIED reference engine (oracle/_ref).

Each fixture holds the flat network, the state after swmm_start, the external inflow hydrographs,
and the reference's own trajectory: simulated time and Picard iteration count after EVERY step
plus full snapshots of the state at ~24 steps.  Run from the repo root after `make -C oracle ref`:
    python tests/golden/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402

CASES = ["c1_tree", "c1_tree_slot", "c2_grid12_slot", "c2_grid12_extran"]


def make(name: str, n_snaps: int = 24):
    e, _ = pc.open_reference(pc.case_inp(name))
    net = e.network()
    s0 = pc.grab_state(e)
    inf = e.inflows()
    t_end = e.total_duration_s()
    times, iters, snaps, snap_steps = [], [], {f: [] for f in pc.SNAP_FIELDS}, []
    step = 0
    snaps_at = set(list(range(1, 12)) + [15, 20, 30, 50, 75, 100, 150, 200, 300, 500, 750, 1000, 1500,
                                          2000, 3000, 4000])
    while True:
        t = e.step()
        step += 1
        times.append(e.routing_time_ms() / 1000.0)
        iters.append(e.last_iterations())
        if step in snaps_at or t == 0:
            snap_steps.append(step)
            for f in pc.SNAP_FIELDS:
                try:
                    snaps[f].append(e.field(f).copy())
                except KeyError:
                    pass
        if t == 0:
            break
    e.end()
    e.close()
    extra = {"s0_" + k: v for k, v in s0.items()}
    extra.update(inf_node=inf["node"], inf_ts_start=inf["ts_start"], inf_ts_t=inf["ts_t"],
                 inf_ts_q=inf["ts_q"], inf_sfactor=inf["sfactor"], inf_baseline=inf["baseline"],
                 inf_concen=inf["concen"] if inf["concen"] is not None else np.zeros(0),
                 inf_start=np.array([inf["start_day"], inf["start_secs"]]), t_end=np.array(t_end),
                 series_time=np.array(times), series_iters=np.array(iters, dtype=np.int32),
                 snap_steps=np.array(snap_steps, dtype=np.int32))
    for f, v in snaps.items():
        if v:
            extra["snap_" + f] = np.array(v)
    net.save(pc.golden_path(name), **extra)
    print(name, "steps", step, "snapshots", len(snap_steps), os.path.getsize(pc.golden_path(name)), "bytes")


if __name__ == "__main__":
    for c in (sys.argv[1:] or CASES):
        make(c)
