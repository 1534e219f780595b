"""Cold path of config 3: topology creation, first solve (device setup), second solve."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.lattice(400, 250); v, p, s = w.prepare()
x0 = v[0][w.free_vars]
# touch the device first so that context creation is not charged to the solver
small = wl.truss(4); sv, sp_, _ = small.prepare()
fk.Topology.from_arrays(small.n_vars, small.kind, small.idx, small.free_vars, small.rows).batch_solve(sv, sp_)
t0 = time.perf_counter()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
t1 = time.perf_counter()
x, r = topo.lm_solve(v[0], p[0], x0)
t2 = time.perf_counter()
x, r = topo.lm_solve(v[0], p[0], x0)
t3 = time.perf_counter()
print("topology %.3f s, first solve %.3f s, second solve %.3f s" % (t1 - t0, t2 - t1, t3 - t2))
