#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU: the 1000 x 500 looped grid (~1M conduits, SLOT, no pollutants) as a
single model (M = 1), stepped entirely on the device: the single-GPU number next to the reference's
4.5e6 conduit-updates/s (8 threads).  tools/c5_partitioned.py runs the same model striped over N GPUs."""
import sys
import time

sys.path.insert(0, ".")
import swmm_b200  # noqa: F401
from swmm_b200 import network, scenarios, solver

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
sim_s = float(sys.argv[3]) if len(sys.argv) > 3 else 600.0
t0 = time.perf_counter()
spec = scenarios.GridSpec(nx=nx, ny=ny, hours=1.0, pollutants=False, surcharge="SLOT")
case = network.build_grid(spec)
print(f"built {case.net.n_links} links / {case.net.n_nodes} nodes in {time.perf_counter() - t0:.1f} s", flush=True)
s = solver.Solver(case.net, 1)
s.load_state(case.state0)
s.set_inflows(**case.inflows)
s.run_steps(5, sim_s)
s.sync()
cu0 = s.conduit_updates()
s.phase_times()
t0 = time.perf_counter()
s.run_steps(1000000, sim_s)
s.sync()
dt = time.perf_counter() - t0
st = s.stats()[0]
cu = s.conduit_updates() - cu0
print(f"sim {st.sim_time:.1f} s, steps {st.steps}, iterations {st.iterations}, wall {dt:.3f} s -> "
      f"{cu / dt:.3e} conduit-updates/s, {200.0 * cu / dt / 1e9:.1f} GB/s algorithmic", s.phase_times())
