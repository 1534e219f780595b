"""A/B of the two batched LM kernels (tile vs sketch-per-thread), device resident: M sketches/s per workload."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

def run(name, w, reps=5):
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    info = topo.sketch_kernel_info()
    plan = topo.plan(w.n)
    plan.upload(v, p)
    fk.lib().fk_batch_plan_sync(plan._h)
    out = {}
    for kernel in ("tile", "sketch_solo", "sketch_pair"):
        if kernel != "tile" and not info["available"]:
            continue
        with api.lm_kernel(kernel):
            for _ in range(2):
                plan.run()
            fk.lib().fk_batch_plan_sync(plan._h)
            best = 1e9
            for _ in range(reps):
                t0 = time.perf_counter(); plan.run(); fk.lib().fk_batch_plan_sync(plan._h); best = min(best, time.perf_counter() - t0)
            x = np.zeros((w.n, topo.info["n_free"])); rep = np.zeros(w.n, dtype=fk.REPORT_DTYPE)
            plan.download(x, rep); fk.lib().fk_batch_plan_sync(plan._h)
            out[kernel] = (best, x, rep)
    line = f"{name:24s} n={w.n:8d} state={info['state_doubles']:4d}"
    for k, (t, x, rep) in out.items():
        line += f"  {k}: {w.n / t / 1e6:8.2f} M/s ({t * 1e3:7.3f} ms)"
    if len(out) == 3:
        same = np.array_equal(out["tile"][2]["trace_hash"], out["sketch_pair"][2]["trace_hash"]) and np.array_equal(out["tile"][2]["trace_hash"], out["sketch_solo"][2]["trace_hash"])
        line += f"  traces equal={same} max|dx| solo {np.max(np.abs(out['tile'][1] - out['sketch_solo'][1])):.1e} pair {np.max(np.abs(out['tile'][1] - out['sketch_pair'][1])):.1e}"
    print(line, flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    run("truss20", wl.truss(n))
    run("cad_mix", wl.cad_mix(n))
    run("truss10", wl.truss(n, n_points=10))
    run("truss14", wl.truss(n, n_points=14))
    run("hinged4", wl.hinged_triangles(4, n))
    for m in (4096, 16384, 262144):
        run("truss20", wl.truss(m))
        run("cad_mix", wl.cad_mix(m))
