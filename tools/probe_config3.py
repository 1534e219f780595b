import os, sys, time; sys.path.insert(0,'.')
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w=wl.lattice(400,250); v,p,s=w.prepare()
x0=v[0][w.free_vars]
for tm,tw in ((8,32768),(16,16384)):
    os.environ['FK_TEAM_MAX']=str(tm); os.environ['FK_TEAM_WORK']=str(tw)
    topo=fk.Topology.from_arrays(w.n_vars,w.kind,w.idx,w.free_vars,w.rows)
    topo.lm_solve(v[0],p[0],x0)
    t0=time.time(); x,r=topo.lm_solve(v[0],p[0],x0); dt=time.time()-t0
    print('team',tm,tw,'solve wall %.3fs'%dt, r['factorizations'], r['ssr'], {k:round(val,2) for k,val in topo.last_timing().items()})
    topo.close()
