import sys, os, tempfile, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import swmm_b200
from swmm_b200 import scenarios, abi, solver
import refengine
EMUL='/root/repo/tests/emul/libswb_emul.so'
n = int(sys.argv[1]); sur = sys.argv[2]; hours=float(sys.argv[3])
d = tempfile.mkdtemp()
spec = scenarios.GridSpec(nx=n, ny=n, hours=hours, surcharge=sur)
open(d+'/c2.inp','w').write(scenarios.c2_grid_inp(spec))
e = refengine.RefEngine()
e.open(d+'/c2.inp'); e.start()
net = e.network()
s = solver.Solver(net, 1, lib_path=EMUL)
state = {}
for f in solver.Solver.STATE_FIELDS:
    try: state[f] = e.field(f)
    except KeyError: pass
s.load_state(state)
s.set_inflows(**e.inflows())
nsteps=0; mx={}
t0=time.time(); flagged=False
F=['SWB_NODE_NEW_DEPTH','SWB_LINK_NEW_FLOW','SWB_NODE_NEW_QUAL','SWB_LINK_NEW_QUAL','SWB_NODE_NEW_VOLUME','SWB_NODE_OVERFLOW']
tend = e.total_duration_s()
while True:
    t = e.step(); nsteps+=1
    s.run_steps(1, tend)
    if t==0: break
    if nsteps%10 and nsteps>20: continue
    for f in F:
        r=e.field(f); m=s.get_field(f)[0]
        err=np.abs(r-m).max(); mx[f]=max(mx.get(f,0),err)
        if err>1e-9 and not flagged:
            flagged=True; k=int(np.abs(r-m).argmax()); print('first mismatch step',nsteps,f,k,r[k],m[k])
st=s.stats()[0]
print('steps', nsteps, 'mine steps', st.steps, 'iters', st.iterations, 'nonconv', st.non_converged, 'ref nonconv', e.non_converge_count(), 'time', time.time()-t0)
print(mx)
