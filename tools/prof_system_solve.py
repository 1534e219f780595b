"""fk_batch_system_solve on the truss batch (for an ncu launch list): prof_system_solve.py [n]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
w = wl.truss(n)
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
for _ in range(3):
    t0 = time.perf_counter()
    x, s, rep = topo.batch_system_solve(w.raw_vars, w.raw_param[0], shared_param=True)
    print("system solve", (time.perf_counter() - t0) * 1e3, "ms")
