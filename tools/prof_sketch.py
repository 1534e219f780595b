"""One launch of the chosen batched LM kernel on a workload (for ncu): prof_sketch.py <tile|sketch> <truss|cad> [n]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl
kernel, what = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
w = wl.truss(n) if what == "truss" else wl.cad_mix(n)
v, p, scale = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
plan = topo.plan(n)
plan.upload(v, p)
with api.lm_kernel(kernel):
    plan.run()
    fk.lib().fk_batch_plan_sync(plan._h)
print("done")
