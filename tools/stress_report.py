"""Config 5 report: exit-reason histogram, trace-equality rate vs the CPU oracle and throughput per
stress family (8,192 sketches each).  Run on a GPU box: python tools/stress_report.py [n_each]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import fiksi_b200 as fk
import oracle
from fiksi_b200 import workloads as wl

n_each = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rows = []
for name, w in wl.stress_families(n_each):
    v, p, scale = w.prepare(perturb=(name != "nan_coincident_points"))
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    topo.batch_solve(v, p)  # warm-up at full size (staging pipeline allocation)
    t0 = time.perf_counter()
    x, rep = topo.batch_solve(v, p)
    dt = time.perf_counter() - t0
    sub = min(n_each, 1024)
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    xo, ro, secs = oracle.lm_solve_batch_uniform(op, v[:sub], p[:sub], threads=os.cpu_count() or 1)
    same = (rep["trace_hash"][:sub] == ro["trace_hash"]) & (rep["exit_reason"][:sub] == ro["exit_reason"])
    err = 0.0
    if same.any() and name != "nan_coincident_points":
        err = float(np.max(np.max(np.abs(x[:sub][same] - xo[same]), axis=1) / np.max(np.abs(xo[same]), axis=1)))
    flagged = ~topo.batch_analyze(w.raw_vars, w.raw_param)          # System::analyze: dependent expressions
    ref_flag = ~np.array([oracle.analyze(w.raw_vars[k], w.kind, w.idx, w.raw_param[k]) for k in range(min(sub, 64))])
    rows.append({"family": name, "analyze_overconstrained_frac": float(flagged.any(axis=1).mean()),
                 "analyze_equal_oracle": bool(np.array_equal(flagged[:len(ref_flag)], ref_flag)), "n": n_each, "tile": topo.info["tile"], "exit_hist": np.bincount(rep["exit_reason"], minlength=5).tolist(),
                 "mean_factorizations": float(rep["factorizations"].mean()), "gpu_e2e_sketches_per_s": n_each / dt,
                 "cpu_port_sketches_per_s": sub / secs, "trace_equal_frac": float(same.mean()), "max_rel_coord_err_where_equal": err})
    print(json.dumps(rows[-1]))
