import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, fiksi_b200 as fk
from fiksi_b200 import workloads as wl
# large path on a lattice big enough to have big fronts (f > 32, ns > 64 unlikely but multi-level)
w = wl.lattice(60, 40); v, p, s = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
print(topo.info["path"], topo.info["n_free"])
x, r = topo.lm_solve(v[0], p[0], v[0][w.free_vars]); print(r)
# batched path + eval kernels, ragged sizes
for maker in (wl.truss, wl.cad_mix):
    w = maker(333); v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x, rep = topo.batch_solve(v, p)
    plan = topo.plan(w.n); plan.upload(v, p)
    m, jn = topo.info["n_rows"], topo.info["jac_nnz"]
    rr = np.zeros((w.n, m)); jj = np.zeros((w.n, jn))
    plan.eval(0); plan.eval_download(rr, jj); plan.eval(1); plan.eval_download(rr, None)
    fk.lib().fk_batch_plan_sync(plan._h)
    print(maker.__name__, np.bincount(rep["exit_reason"]))
