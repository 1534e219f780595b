import sys, time; sys.path.insert(0,'.')
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
t0=time.time(); w=wl.lattice(400,250); v,p,s=w.prepare(); print('gen+prepare %.2fs'%(time.time()-t0))
t0=time.time(); topo=fk.Topology.from_arrays(w.n_vars,w.kind,w.idx,w.free_vars,w.rows); print('symbolic %.2fs'%(time.time()-t0), topo.info)
x0=v[0][w.free_vars]
for rep in range(3):
    t0=time.time(); x,r=topo.lm_solve(v[0],p[0],x0); dt=time.time()-t0
    print('solve wall %.3fs'%dt, r, topo.last_timing())
rr,jj,ms=topo.eval_large(v[0],p[0],x0,repeats=20,want_j=False)
print('eval ms',ms,'GB/s alg', topo.info['eval_bytes']/ms/1e6)
print('fp64 peak', fk.fp64_peak_tflops())
tm=topo.last_timing(); fl=topo.info['chol_flops']
print('factor TFLOP/s', fl*tm['factors']/(tm['factor_ms']*1e-3)/1e12)
