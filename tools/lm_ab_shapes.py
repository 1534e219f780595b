"""A/B of the sketch-kernel shapes on one workload, device resident, optionally against another build of the library:
   FK_AB_LIB=/path/to/other/libfiksi_b200.so python tools/lm_ab_shapes.py truss20 65536 sketch_solo sketch_pair sketch_quad
Prints M sketches/s per shape and whether every report field and coordinate equals the first shape's."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fiksi_b200._lib as _lib
if os.environ.get("FK_AB_LIB"):
    _lib.LIB_PATH = os.environ["FK_AB_LIB"]
import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

WORKLOADS = {"truss20": lambda n: wl.truss(n), "truss14": lambda n: wl.truss(n, n_points=14), "truss10": lambda n: wl.truss(n, n_points=10),
             "cad_mix": lambda n: wl.cad_mix(n), "hinged4": lambda n: wl.hinged_triangles(4, n), "hinged16": lambda n: wl.hinged_triangles(16, n)}

def main():
    name, n, shapes = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    w = WORKLOADS[name](n)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    plan = topo.plan(w.n)
    plan.upload(v, p)
    fk.lib().fk_batch_plan_sync(plan._h)
    first = None
    line = f"{os.path.basename(os.path.dirname(_lib.LIB_PATH))}/{os.path.basename(_lib.LIB_PATH)} {name} n={n}"
    for shape in shapes:
        try:
            ctx = api.lm_kernel(shape)
        except Exception as e:  # an older build without that shape
            line += f"  {shape}: n/a"
            continue
        with ctx:
            for _ in range(3):
                plan.run()
            fk.lib().fk_batch_plan_sync(plan._h)
            ts = []
            for _ in range(9):
                t0 = time.perf_counter(); plan.run(); fk.lib().fk_batch_plan_sync(plan._h); ts.append(time.perf_counter() - t0)
            t = float(np.median(ts))
            x = np.zeros((w.n, topo.info["n_free"])); rep = np.zeros(w.n, dtype=fk.REPORT_DTYPE)
            plan.download(x, rep); fk.lib().fk_batch_plan_sync(plan._h)
        same = ""
        if first is None:
            first = (x, rep)
        else:
            same = " same" if (np.array_equal(x, first[0], equal_nan=True) and all(np.array_equal(rep[k], first[1][k], equal_nan=True) for k in rep.dtype.names)) else " DIFFERENT"
        line += f"  {shape}: {w.n / t / 1e6:7.2f} M/s ({t * 1e3:6.3f} ms){same}"
    print(line, flush=True)

if __name__ == "__main__":
    main()
