"""Soak of the polling kernels (tile dataflow factorisation, chained solves): N LM solves of config 3, every one
bit-compared with the first; also a mid-size irregular system.  A hang or a timing-dependent sum would show here."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for name, w in (("lattice 400x250", wl.lattice(400, 250)), ("lattice 150x120", wl.lattice(150, 120)), ("lattice 61x47", wl.lattice(61, 47))):
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x0 = v[0][w.free_vars]
    ref, rep0 = topo.lm_solve(v[0], p[0], x0)
    t0 = time.perf_counter()
    bad = 0
    for k in range(n):
        x, rep = topo.lm_solve(v[0], p[0], x0)
        bad += int(not np.array_equal(x, ref) or rep["trace_hash"] != rep0["trace_hash"])
    print(f"{name}: {n} solves, {1e3 * (time.perf_counter() - t0) / n:.2f} ms each, mismatches {bad}, exit {rep0['exit_reason']}")
