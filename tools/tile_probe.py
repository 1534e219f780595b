"""LM kernel throughput vs lanes per sketch for trusses of different size (device-resident)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
npts = int(sys.argv[1]); n = 65536
w = wl.truss(n, n_points=npts); v, p, s = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
plan = topo.plan(n); st = torch.cuda.current_stream().cuda_stream
plan.upload(v, p, st); plan.run(st); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): plan.run(st)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(json.dumps({"points": npts, "tile": topo.info["tile"], "smem_per_sketch": topo.info["smem_bytes"], "Msketches_per_s": round(n / ms / 1e3, 2)}))
