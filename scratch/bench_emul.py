import sys, runpy
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import parity_common as pc
from swmm_b200 import solver
solver.CUDA_LIB = pc.EMUL_LIB
sys.argv = ['bench.py','--grid','8','--members','32','--steps','3','--warmup','1','--spinup','600','--routing-steps','5','--cpu-steps','20','--hours','2']
runpy.run_path('/root/repo/bench.py', run_name='__main__')
