"""In-library multi-device path from ONE process (fk_batch_solve(n_gpus=N): one host thread per device, contiguous
ranges, no data-path collective) next to the single-device call, on the truss batch.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

ndev = fk.device_count()
per_gpu = 65536
out = {"devices_visible": ndev, "sketches_per_gpu": per_gpu, "runs": []}
for g in [x for x in (1, 2, 4, 8) if x <= ndev]:
    n = per_gpu * g
    w = wl.truss(n)
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    topo.batch_solve(v, p, n_gpus=g)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); x, rep = topo.batch_solve(v, p, n_gpus=g); best = min(best, time.perf_counter() - t0)
    x1, rep1 = topo.batch_solve(v[:4096], p[:4096], n_gpus=1)
    out["runs"].append({"n_gpus": g, "sketches": n, "sketches_per_s": n / best, "ms": best * 1e3,
                        "equal_to_single_device": bool(np.array_equal(x[:4096], x1) and np.array_equal(rep["trace_hash"][:4096], rep1["trace_hash"])),
                        "host_buffers": "pageable numpy arrays"})
print(json.dumps(out))
