import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import fiksi_b200 as fk
from fiksi_b200 import api
from test_recursive_assembly import _pose_batch
kernel, n = sys.argv[1], int(sys.argv[2])
kind, idx, v, p = _pose_batch(n)
topo = fk.Topology.from_arrays(9, kind, idx, np.arange(7), np.arange(3))
print(topo.info, topo.sketch_kernel_info())
with api.lm_kernel(kernel):
    x, r = topo.batch_solve(v, p)
print(kernel, n, "ok", np.bincount(r["exit_reason"]), float(np.mean(r["ssr"] < 1e-8)))
