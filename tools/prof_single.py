"""One single-system LM solve of the config-1 sketch (for ncu): prof_single.py [hinged triangles]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.hinged_triangles(int(sys.argv[1])) if len(sys.argv) > 1 else wl.cad_mix(1)
v, p, s = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
x0 = v[0][w.free_vars]
for _ in range(4):
    x, r = topo.lm_solve(v[0], p[0], x0)
print(r)
