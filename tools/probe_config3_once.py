import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.lattice(400, 250); v, p, s = w.prepare()
x0 = v[0][w.free_vars]
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
x, r = topo.lm_solve(v[0], p[0], x0)
print(r, topo.last_timing())
