import sys, time
sys.path.insert(0, ".")
import swmm_b200
from swmm_b200 import network, scenarios, solver
case = network.build_grid(scenarios.GridSpec(nx=100, ny=100, hours=2.0))
s = solver.Solver(case.net, 1)
s.load_state(case.state0); s.set_inflows(**case.inflows)
s.run_steps(700, case.t_end)
s.run_steps(300, case.t_end)
print(s.stats()[0].steps, s.phase_times())
