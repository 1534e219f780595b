"""K1/K2 timing on its own: config-4 topology (1M sketches) and the truss (65,536 x 16)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

def run(name, w, mode):
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    plan = topo.plan(w.n)
    stream = torch.cuda.current_stream().cuda_stream
    plan.upload(v, p, stream)
    torch.cuda.synchronize()
    for _ in range(3):
        plan.eval(mode, stream)
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(mode, stream); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    i = topo.info
    alg = i["eval_bytes"] * w.n
    dram = 8 * w.n * (i["n_vars"] + i["n_expr"] + i["n_rows"] + (i["jac_nnz"] if mode == 0 else 0))
    print(json.dumps({"workload": name, "mode": mode, "n": w.n, "ms": ms, "alg_gbs": alg / ms / 1e6, "min_dram_gbs": dram / ms / 1e6}))
    plan.close()

if __name__ == "__main__":
    run("cad_mix", wl.cad_mix(1_000_000), 0)
    run("cad_mix", wl.cad_mix(1_000_000), 1)
    run("truss", wl.truss(262144), 0)
