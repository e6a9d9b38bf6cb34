#!/usr/bin/env python
"""Writes tests/golden/c5_grid_1000x500.npz from the UNMODIFIED reference engine (oracle/_ref): BASELINE config 5,
the 1000 x 500 looped grid (998 501 conduits, 500 001 nodes, SLOT, no pollutants), 10 simulated minutes.

The full state of a 1M-link model is too large for a fixture, so it holds the reference's simulated time and
Picard count after EVERY step and, at a few steps, (a) every `STRIDE`-th node depth / node volume / link flow /
link depth and (b) the sum and the sum of squares of the COMPLETE arrays (a checksum that every object enters).
About 6-10 minutes of CPU (most of it parsing the 100 MB .inp):
    python tests/golden/make_c5_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402
from swmm_b200 import scenarios  # noqa: E402

NX, NY, SIM_MIN, STRIDE = 1000, 500, 10.0, 97
FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_LINK_NEW_FLOW", "SWB_LINK_NEW_DEPTH"]
SNAPS_AT = {1, 2, 5, 10, 20, 40, 80}


def main():
    t0 = time.time()
    spec = scenarios.GridSpec(nx=NX, ny=NY, hours=SIM_MIN / 60.0, pollutants=False, surcharge="SLOT", threads=os.cpu_count())
    inp = scenarios.c2_grid_inp(spec)
    print(f".inp written to memory: {len(inp) / 1e6:.0f} MB in {time.time() - t0:.0f} s", flush=True)
    e, _ = pc.open_reference(inp)
    print(f"reference opened and started after {time.time() - t0:.0f} s", flush=True)
    out = {"nx": NX, "ny": NY, "sim_min": SIM_MIN, "stride": STRIDE}
    times, iters, snap_steps = [], [], []
    samples = {f: [] for f in FIELDS}
    sums = {f: [] for f in FIELDS}
    step = 0
    try:
        while True:
            t = e.step()
            step += 1
            times.append(e.routing_time_ms() / 1000.0)
            iters.append(e.last_iterations())
            if step in SNAPS_AT or t == 0:
                snap_steps.append(step)
                for f in FIELDS:
                    a = e.field(f)
                    samples[f].append(a[::STRIDE].copy())
                    sums[f].append([float(np.sum(a)), float(np.sum(a * a))])
            if t == 0:
                break
    finally:
        e.end()
        e.close()
    out.update(series_time=np.array(times), series_iters=np.array(iters, dtype=np.int32),
               snap_steps=np.array(snap_steps, dtype=np.int32))
    for f in FIELDS:
        out["sample_" + f] = np.array(samples[f])
        out["sums_" + f] = np.array(sums[f])
    path = os.path.join(HERE, f"c5_grid_{NX}x{NY}.npz")
    np.savez_compressed(path, **out)
    print("steps", step, "iterations", int(np.sum(iters)), "snapshots", snap_steps, os.path.getsize(path), "bytes,",
          f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
