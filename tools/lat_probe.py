import sys, time
sys.path.insert(0, '.')
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
def lat(w, reps=100):
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x0 = v[0][w.free_vars]
    for _ in range(5): topo.lm_solve(v[0], p[0], x0)
    t0 = time.perf_counter()
    for _ in range(reps): x, r = topo.lm_solve(v[0], p[0], x0)
    return (time.perf_counter() - t0) / reps * 1e6, topo.info["tile"], int(r["factorizations"])
print("cad_mix", lat(wl.cad_mix(1)))
for nt in (1, 4, 16, 64): print("hinged", nt, lat(wl.hinged_triangles(nt)))
print("truss", lat(wl.truss(1)))
