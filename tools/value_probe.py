"""Device-resident rate of the batched LM kernel on the bench workload (65,536 trusses): python tools/value_probe.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 65536
w = wl.truss(n)
v, p, scale = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
plan = topo.plan(n)
stream = torch.cuda.current_stream().cuda_stream
plan.upload(v, p, stream)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3): plan.run(stream)
ts = []
for _ in range(steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run(stream); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sum(ts) / len(ts)
print(f"{n / ms / 1e3:.2f} M sketches/s device resident, {ms:.4f} ms per step (kernel: {topo.batch_kernel(n)})")
