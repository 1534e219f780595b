"""SURVEY H1/H2 hunt: normal-equation LDL^T (product) vs Householder QR (reference algorithm, oracle) on
harder starts — larger noise means more LM iterations and smaller damping.  Prints per noise level the
fraction of sketches with identical decision traces, the largest coordinate difference among those, the
smallest final lambda and the exit histogram."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fiksi_b200 as fk, oracle
from fiksi_b200 import workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for noise in (0.05, 0.5, 1.0, 2.0, 4.0):
    for name, w in (("truss", wl.truss(n, noise=noise)),):
        v, p, s = w.prepare()
        topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
        xg, rg = topo.batch_solve(v, p)
        op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
        xo, ro, _ = oracle.lm_solve_batch_uniform(op, v, p, threads=os.cpu_count() or 1)
        same = (rg["trace_hash"] == ro["trace_hash"]) & (rg["exit_reason"] == ro["exit_reason"])
        err = np.max(np.abs(xg - xo), axis=1) / np.max(np.abs(xo), axis=1)
        print(json.dumps({"workload": name, "noise": noise, "n": n, "trace_equal_frac": float(same.mean()),
                          "max_rel_coord_err_where_equal": float(err[same].max()) if same.any() else None,
                          "max_rel_coord_err_where_different": float(err[~same].max()) if (~same).any() else None,
                          "min_final_lambda": float(ro["lambda"].min()), "max_factorizations": int(ro["factorizations"].max()),
                          "exit_hist_gpu": np.bincount(rg["exit_reason"], minlength=5).tolist(),
                          "exit_hist_cpu": np.bincount(ro["exit_reason"], minlength=5).tolist()}))
