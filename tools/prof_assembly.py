import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.cad_mix(1_000_000)
v, p, scale = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
plan = topo.plan(w.n)
plan.upload(v, p, 0)
torch.cuda.synchronize()
for _ in range(3):
    plan.eval(0, 0)
torch.cuda.synchronize()
