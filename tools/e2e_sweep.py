"""End-to-end rate of fk_batch_system_solve on the bench workload (65,536 trusses, pinned host buffers) for the chunk
count given in FK_E2E_CHUNKS (read once by the library): python tools/e2e_sweep.py [steps]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 65536
w = wl.truss(n)
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
info = topo.info
hraw = torch.from_numpy(w.raw_vars).pin_memory()
hrawp = torch.from_numpy(np.ascontiguousarray(w.raw_param[0])).pin_memory()
hout = torch.empty((n, info["n_free"]), dtype=torch.float64).pin_memory()
hrep = torch.empty((n, 5), dtype=torch.float64).pin_memory()
def step():
    topo.batch_system_solve_into(0, n, hraw.data_ptr(), hrawp.data_ptr(), hout.data_ptr(), hrep.data_ptr(), shared_param=True)
stream = os.environ.get("FK_SWEEP_STREAM") == "1"  # steps streamed two deep through the two halves of the call
hout2, hrep2 = torch.empty_like(hout).pin_memory(), torch.empty_like(hrep).pin_memory()
outs = ((hout, hrep), (hout2, hrep2))
def begin(i):
    return topo.batch_system_solve_begin(0, n, hraw.data_ptr(), hrawp.data_ptr(), outs[i & 1][0].data_ptr(), outs[i & 1][1].data_ptr(), shared_param=True)
for _ in range(3): step()
t0 = time.perf_counter()
if stream:
    prev = begin(0)
    for i in range(1, steps):
        cur = begin(i)
        topo.batch_system_solve_wait(prev)
        prev = cur
    topo.batch_system_solve_wait(prev)
else:
    for _ in range(steps): step()
dt = time.perf_counter() - t0
print(("streamed " if stream else "") + f"FK_E2E_CHUNKS={os.environ.get('FK_E2E_CHUNKS', '8 (default)')}: {n * steps / dt / 1e6:.2f} M sketches/s end to end, {dt / steps * 1e6:.0f} us per call")
