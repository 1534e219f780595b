import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 250)
w = wl.lattice(nx, ny); v, p, s = w.prepare()
x0 = v[0][w.free_vars]
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
topo.lm_solve(v[0], p[0], x0)
for rep in range(2):
    t = time.time(); x, r = topo.lm_solve(v[0], p[0], x0); dt = time.time() - t
    print(f"solve wall {dt:.3f}s", r["factorizations"], r["ssr"], topo.last_timing())
