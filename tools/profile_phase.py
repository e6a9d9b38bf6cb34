#!/usr/bin/env python
"""One Picard phase in isolation for ncu (the persistent kernel otherwise mixes all phases in one launch).

    ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/links \\
        python tools/profile_phase.py links [--members 512] [--steps 10] [--lib vlib/x.so]

Spins the bench ensemble up, then runs ONE launch of `--steps` routing steps of dynwave_execute with the
other Picard phase switched off, bracketed by cudaProfilerStart/Stop.  Prints the kernel time."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swmm_b200  # noqa: E402,F401
from swmm_b200 import solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("phase", choices=["links", "nodes", "both"])
ap.add_argument("--members", type=int, default=512)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--spinup", type=float, default=6000.0)
ap.add_argument("--lib", default=None)
a = ap.parse_args()
if a.lib:
    solver.CUDA_LIB = os.path.abspath(a.lib)
import bench  # noqa: E402

args = argparse.Namespace(grid=100, hours=6.0, surcharge="SLOT", members=a.members, members_total=a.members)
s, case, spec = bench.make_ensemble(args, 0, 0)
bench.spin_up(s, a.spinup)
PH_DYNWAVE = 4
dbg = {"links": 2, "nodes": 1, "both": 0}[a.phase]
s.debug_run(PH_DYNWAVE, a.steps, dbg, False)          # warm (instruction cache, L2)
s.phase_times(reset=True)
s.debug_run(PH_DYNWAVE, a.steps, dbg, True)
print(a.phase, "kernel ms", s.last_kernel_ms(), {k: round(v, 2) for k, v in s.phase_times().items() if v})
s.close()
