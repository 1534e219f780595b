"""One launch each of the L-BFGS and analyze kernels on the truss batch (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.truss(65536); v, p, s = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
plan = topo.plan(w.n); plan.upload(v, p, 0)
plan.run_lbfgs(0); plan.run_lbfgs(0)
torch.cuda.synchronize()
topo.batch_analyze(w.raw_vars[:16384], w.raw_param[:16384]); topo.batch_analyze(w.raw_vars[:16384], w.raw_param[:16384])
