import sys, os, tempfile, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import swmm_b200
from swmm_b200 import scenarios, abi, solver
import refengine
EMUL='/root/repo/tests/emul/libswb_emul.so'
d = tempfile.mkdtemp()
spec = scenarios.TreeSpec()
open(d+'/c1.inp','w').write(scenarios.c1_tree_inp(spec))
e = refengine.RefEngine()
e.open(d+'/c1.inp'); e.start()
net = e.network()
s = solver.Solver(net, 1, lib_path=EMUL)
state = {}
for f in solver.Solver.STATE_FIELDS:
    try: state[f] = e.field(f)
    except KeyError: pass
s.load_state(state)
s.set_inflows(**e.inflows())
nsteps=0; maxd=0; maxq=0
t0=time.time()
while True:
    t = e.step(); nsteps+=1
    s.run_steps(1, spec.hours*3600)
    if t==0: break
    dr = e.field('SWB_NODE_NEW_DEPTH'); dm = s.get_field('SWB_NODE_NEW_DEPTH')[0]
    qr = e.field('SWB_LINK_NEW_FLOW'); qm = s.get_field('SWB_LINK_NEW_FLOW')[0]
    ed = np.abs(dr-dm).max(); eq = np.abs(qr-qm).max()
    maxd=max(maxd,ed); maxq=max(maxq,eq)
    if nsteps<4 or (max(ed,eq)>1e-6 and not globals().get("flagged")):
        flagged = max(ed,eq)>1e-6
        st = s.stats()[0]
        print(nsteps, 't_ref(ms)', e.routing_time_ms(), 't_mine', st.sim_time, 'dt', st.last_dt, 'iters', st.iterations, 'maxerr d,q', ed, eq)
print('steps', nsteps, 'max abs err depth', maxd, 'flow', maxq, 'time', time.time()-t0)
