import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, oracle, fiksi_b200 as fk, scenarios as sc
for name in ("single_triangle", "collinear_points"):
    b = sc.ALL[name](oracle.System)
    for prob, scale, keep in b["s"].prepare(perturb=True):
        vars_, kind, idx, param, free_vars, rows = keep
        topo = fk.Topology.from_arrays(len(vars_), kind, idx, free_vars, rows)
        xo, ro = oracle.lbfgs_solve(prob, vars_[free_vars])
        print(name, 'cpu', ro['lambda'].hex(), ro['factorizations'])
        for rep in range(3):
            xg, rg = topo.batch_solve_lbfgs(vars_[None,:], np.asarray(param)[None,:])
            print('   gpu', float(rg['lambda'][0]).hex(), rg['factorizations'][0], rg['exit_reason'][0])
        # batch of identical sketches: all tiles of a warp
        V = np.repeat(vars_[None,:], 64, axis=0); Pm = np.repeat(np.asarray(param)[None,:], 64, axis=0)
        xg, rg = topo.batch_solve_lbfgs(V, Pm)
        print('   gpu x64 distinct lambdas', sorted(set(float(v).hex() for v in rg['lambda'])), sorted(set(rg['factorizations'].tolist())))
