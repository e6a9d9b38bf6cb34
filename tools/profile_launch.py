#!/usr/bin/env python
"""The bench's timed launch, bracketed for ncu (`--profile-from-start off`): the C4 ensemble is built and
spun up exactly as bench.py does, one warm launch, then ONE swb_run_steps launch of `--routing-steps`
routing steps between cuProfilerStart/Stop.  Writes the launch's own conduit-update count and event time
to `--out` so traffic / algorithmic bytes is like for like (profiles/ncu_traffic_r02.json).

    ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/x \\
        python tools/profile_launch.py --members 4096 --out gpurun_out/x_launch.json"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swmm_b200  # noqa: E402,F401
from swmm_b200 import solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--routing-steps", type=int, default=10)
ap.add_argument("--spinup", type=float, default=6000.0)
ap.add_argument("--lib", default=None)
ap.add_argument("--out", default=None)
a = ap.parse_args()
if a.lib:
    solver.CUDA_LIB = os.path.abspath(a.lib)
import bench  # noqa: E402

args = argparse.Namespace(grid=100, hours=6.0, surcharge="SLOT", members=a.members, members_total=a.members,
                          member_order="scale")
s, case, spec = bench.make_ensemble(args, 0, 0)
bench.spin_up(s, a.spinup, reorder_every=100)       # the bench's own configuration
s.run_steps(a.routing_steps, case.t_end)
s.sync()
cu0 = s.conduit_updates()
s.phase_times(reset=True)
cuda = ctypes.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
s.run_steps(a.routing_steps, case.t_end)
s.sync()
cuda.cuProfilerStop()
rec = {"members": a.members, "grid": 100, "routing_steps": a.routing_steps,
       "conduit_updates_in_launch": s.conduit_updates() - cu0, "event_ms": s.last_kernel_ms(),
       "phase_ms": {k: round(v, 3) for k, v in s.phase_times().items() if v}}
print(json.dumps(rec))
if a.out:
    json.dump(rec, open(a.out, "w"))
s.close()
