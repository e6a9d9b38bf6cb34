import sys, time, numpy as np
sys.path.insert(0, ".")
import swmm_b200
from swmm_b200 import network, scenarios, solver
case = network.build_grid(scenarios.GridSpec(nx=100, ny=100, hours=2.0))
s = solver.Solver(case.net, 1)
s.load_state(case.state0); s.set_inflows(**case.inflows)
t0 = time.perf_counter(); s.run_steps(100000, case.t_end); s.sync(); dt = time.perf_counter() - t0
st = s.stats()[0]
print("single 10k-node model, 2 h: steps", st.steps, "iterations", st.iterations, "wall", round(dt, 2), "s ->", s.conduit_updates() / dt, "cu/s", "phases", s.phase_times())
