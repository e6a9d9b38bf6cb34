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
def grab():
    st={}
    for f in solver.Solver.STATE_FIELDS:
        try: st[f] = e.field(f)
        except KeyError: pass
    return st
s.load_state(grab())
leaves = spec.leaves(); node=[i-1 for i in leaves]; n=len(node)
ts_t = np.tile(np.array([0,2,4,6.0])*3600, n); ts_q = np.tile(np.array([0,spec.peak_cfs,0,0.0]), n)
sf = [float(f"{0.5 + 0.5*((k*7919)%100)/100.0:.4f}") for k in range(n)]
s.set_inflows(node, np.arange(n+1)*4, ts_t, ts_q, sf, np.zeros(n))
for step in range(1,12):
    e.step(); s.run_steps(1, spec.hours*3600)
    bad=False
    for f in ['SWB_NODE_NEW_LATFLOW','SWB_NODE_NEW_DEPTH','SWB_NODE_INFLOW','SWB_NODE_OUTFLOW','SWB_NODE_NEW_VOLUME','SWB_LINK_NEW_FLOW','SWB_LINK_NEW_DEPTH','SWB_LINK_NEW_VOLUME','SWB_LINK_DQDH','SWB_LINK_FROUDE','SWB_LINK_FLOW_CLASS','SWB_LINK_SURF_AREA1','SWB_LINK_SURF_AREA2','SWB_COND_A1','SWB_COND_Q1','SWB_NODE_OLD_NET_INFLOW']:
        r=e.field(f); m=s.get_field(f)[0]
        err=np.abs(r-m); k=int(err.argmax())
        if err[k]>1e-9:
            bad=True
            print(step,f,'idx',k,'ref',r[k],'mine',m[k])
    if bad: break
    # resync my state to the reference's to isolate per-step errors
