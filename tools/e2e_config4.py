"""End-to-end rate of fk_batch_solve_device on config 4 (1,000,000 mixed-primitive sketches, pinned host buffers) for the chunk count
in FK_E2E_CHUNKS (read once by the library): python tools/e2e_config4.py [steps]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = 1_000_000
w = wl.cad_mix(n)
v, p, scale = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
hv, hp = torch.from_numpy(v).pin_memory(), torch.from_numpy(p).pin_memory()
hout = torch.empty((n, topo.info["n_free"]), dtype=torch.float64).pin_memory()
hrep = torch.empty((n, 5), dtype=torch.float64).pin_memory()
def step():
    topo.batch_solve_into(0, n, hv.data_ptr(), hp.data_ptr(), hout.data_ptr(), hrep.data_ptr())
for _ in range(2): step()
t0 = time.perf_counter()
for _ in range(steps): step()
dt = time.perf_counter() - t0
print(f"FK_E2E_CHUNKS={os.environ.get('FK_E2E_CHUNKS', '8 (default)')}: {n * steps / dt / 1e6:.1f} M sketches/s end to end, {dt / steps * 1e3:.3f} ms per call")
