"""Heterogeneous batch through fk_lm_solve_batch: many different small topologies, a few members each."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

def build(n_topo, members):
    probs, x0s, keep = [], [], []
    rng = np.random.default_rng(0)
    for t in range(n_topo):
        n_points = 4 + t % 20
        w = wl.truss(members, n_points=n_points, seed=0xF1C50002 + 1000 * t)
        # make the topology unique: fix a different subset of coordinates
        fixed = set(rng.choice(2 * n_points, size=t // 20 % 3, replace=False).tolist()) if t >= 20 else set()
        free = np.array([q for q in range(2 * n_points) if q not in fixed], np.uint32)
        v, p, s = w.prepare()
        for j in range(members):
            fp, k = fk.make_problem(v[j], w.kind, w.idx, p[j], free, w.rows)
            probs.append(fp); keep.append(k); x0s.append(v[j][free])
    return probs, x0s, keep

for n_topo, members in ((50, 20), (200, 5), (1000, 1)):
    probs, x0s, keep = build(n_topo, members)
    api.topology_cache_clear()
    t0 = time.perf_counter(); xs, reps = fk.lm_solve_batch(probs, x0s); cold = time.perf_counter() - t0
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); xs, reps = fk.lm_solve_batch(probs, x0s); best = min(best, time.perf_counter() - t0)
    print(f"{n_topo} topologies x {members}: cold {cold*1e3:.1f} ms, warm {best*1e3:.2f} ms = {len(probs)/best:.0f} systems/s, cache {api.topology_cache_stats()}, converged {np.mean(reps['ssr']<1e-8):.3f}")
