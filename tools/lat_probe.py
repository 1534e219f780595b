"""Single-system latency through fk_topology_lm_solve (cached topology) beside the CPU oracle on one core."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fiksi_b200 as fk, oracle
from fiksi_b200 import workloads as wl
def lat(w, reps=100):
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x0 = v[0][w.free_vars]
    for _ in range(5): topo.lm_solve(v[0], p[0], x0)
    t0 = time.perf_counter()
    for _ in range(reps): x, r = topo.lm_solve(v[0], p[0], x0)
    gpu = (time.perf_counter() - t0) / reps * 1e6
    xb, rb = topo.batch_solve(v[:1], p[:1])
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    creps = max(3, min(100, 20000 // max(1, len(w.free_vars))))
    t0 = time.perf_counter()
    for _ in range(creps): xo, ro, _ = oracle.lm_solve(op, x0)
    cpu = (time.perf_counter() - t0) / creps * 1e6
    return {"gpu_us": round(gpu, 1), "cpu_us": round(cpu, 1), "n": int(len(w.free_vars)), "fact": int(r["factorizations"]),
            "same_as_batch_kernel": bool(np.array_equal(x, xb[0]) and r["trace_hash"] == rb["trace_hash"][0]), "same_trace_as_oracle": bool(r["trace_hash"] == ro["trace_hash"])}
print("cad_mix", lat(wl.cad_mix(1)))
for nt in (1, 4, 16, 64, 128): print("hinged", nt, lat(wl.hinged_triangles(nt)))
print("truss20", lat(wl.truss(1)))
print("truss60", lat(wl.truss(1, n_points=60)))
print("lattice10x8", lat(wl.lattice(10, 8)))
