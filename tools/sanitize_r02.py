"""Small runs of the round-2 kernels for compute-sanitizer (memcheck / racecheck): sketch kernel in both shapes, the
heterogeneous kernel, the device-side pre/post-processing."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

for maker in (lambda: wl.truss(96), lambda: wl.cad_mix(70), lambda: [f for f in wl.stress_families(64) if f[0] == "fixed_point_triangle"][0][1]):
    w = maker()
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    ref = None
    for shape in ("tile", "sketch_solo", "sketch_pair"):
        with api.lm_kernel(shape):
            x, rep = topo.batch_solve(v, p)
        if ref is None:
            ref = (x, rep)
        assert np.array_equal(x, ref[0], equal_nan=True) and np.array_equal(rep["trace_hash"], ref[1]["trace_hash"]), shape
    xs, sc, rp = topo.batch_system_solve(w.raw_vars, w.raw_param, perturb_vars=w.perturb_vars)
probs, x0s, keep = [], [], []
for n_points in range(4, 12):
    w = wl.truss(2, n_points=n_points)
    v, p, s = w.prepare()
    for j in range(2):
        fp, k = fk.make_problem(v[j], w.kind, w.idx, p[j], w.free_vars, w.rows)
        probs.append(fp); keep.append(k); x0s.append(v[j][w.free_vars])
xs, reps = fk.lm_solve_batch(probs, x0s)
print("sanitize run ok", len(probs), np.mean(reps["ssr"] < 1e-8))
