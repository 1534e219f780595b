"""Decomposer::SinglePass.  Host side: the maximum matching / strongly-connected-
expressions plan of fiksi/src/analyze/graph/equations.rs in the product (csrc/single_pass.hpp,
dense-array implementation) vs the oracle's line-by-line restatement (IndexMap semantics), plus the
reference's own test (equations.rs:574-601).  GPU side: fk_system_solve_opts(SinglePass) vs the
oracle's restatement of assemble/mod.rs:169-210.  The reference hands out an SCC's free variables in
HashSet order (unspecified); both sides use ascending order, so the sub-problems are identical."""
import numpy as np
import pytest

import scenarios as sc
import fiksi_b200.system as fsys


def _chain(System, n_tri, rng=None):
    """A chain of triangles hinged at shared points (fiksi/benches/fiksi_bench.rs shape)."""
    s = System()
    pts = [s.add_point(0.0, 0.0), s.add_point(1.0, 0.1)]
    s.point_point_distance(pts[0], pts[1], 1.0)
    for k in range(n_tri):
        p = s.add_point(0.5 + 0.5 * k, 0.9 if k % 2 == 0 else -0.1)
        s.point_point_distance(p, pts[-1], 1.0)
        s.point_point_distance(p, pts[-2], 1.0)
        pts.append(p)
    return s


@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_plan_matches_oracle_on_reference_scenarios(oracle, name):
    plan_o = sc.ALL[name](oracle.System)["s"].single_pass_plan()
    plan_p = sc.ALL[name](fsys.System)["s"].single_pass_plan()
    assert plan_p == plan_o
    for free, exprs in plan_o:
        assert free == sorted(set(free)) and len(exprs) >= 1


def test_plan_properties_on_a_chain(oracle):
    for n_tri in (1, 5, 40):
        so, sp = _chain(oracle.System, n_tri), _chain(fsys.System, n_tri)
        plan = so.single_pass_plan()
        assert sp.single_pass_plan() == plan
        # every expression appears in exactly one step; the free variables of the steps are disjoint
        exprs = [e for _, ex in plan for e in ex]
        assert sorted(exprs) == list(range(1 + 2 * n_tri))
        frees = [v for fv, _ in plan for v in fv]
        assert len(frees) == len(set(frees))
        # a later triangle only needs its own apex: two variables, two expressions
        assert all(len(fv) == 2 and len(ex) == 2 for fv, ex in plan[1:])


def test_matched_and_unmatched_free_variables(oracle):
    """One distance row on a point whose partner is fixed: one of the point's coordinates is matched to
    the expression, the other stays unmatched but free, and both belong to the single step
    (equations.rs:201-213: `var == matched_var || !is_a_matched(var) && free.contains(var)`)."""
    so, sp = oracle.System(), fsys.System()
    for S in (so, sp):
        p0, p1 = S.add_point(0.0, 0.0), S.add_point(1.0, 0.0)
        S.fix(p0)
        S.point_point_distance(p0, p1, 1.0)
    plan = so.single_pass_plan()
    assert sp.single_pass_plan() == plan
    assert len(plan) == 1 and plan[0][1] == [0] and plan[0][0] == [2, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_single_pass_solve_matches_oracle(oracle, name):
    so = sc.ALL[name](oracle.System)["s"]
    sp = sc.ALL[name](fsys.System)["s"]
    so.solve_single_pass()
    sp.solve_single_pass()
    ro, rp = so.reports(), sp.reports()
    assert len(ro) == len(rp)
    for a, b in zip(ro, rp):
        assert a["exit_reason"] == b["exit_reason"] and a["trace_hash"] == b["trace_hash"], (name, a, b)
    vo, vp = np.asarray(so.variables), np.asarray(sp.variables)
    assert np.max(np.abs(vo - vp)) <= 1e-9 * max(np.max(np.abs(vo)), 1e-300)


@pytest.mark.gpu
def test_single_pass_chain_reaches_the_reference_threshold(oracle):
    so, sp = _chain(oracle.System, 16), _chain(fsys.System, 16)
    so.solve_single_pass(); sp.solve_single_pass()
    res = np.abs(sp.residuals())
    assert np.sqrt(np.mean(res ** 2)) < 1e-4                      # fiksi/src/tests/mod.rs:13
    assert np.max(np.abs(np.asarray(so.variables) - np.asarray(sp.variables))) <= 1e-9 * np.max(np.abs(so.variables))
    assert len(sp.reports()) == len(so.reports()) == len(so.single_pass_plan())


@pytest.mark.gpu
@pytest.mark.parametrize("maker", ["truss", "cad_mix", "hinged"])
def test_batched_single_pass_matches_oracle(oracle, maker):
    """fk_batch_solve_single_pass: one batched LM launch per strongly connected set, over all sketches, vs
    the oracle running assemble/mod.rs:169-210 sketch by sketch on the same prepared inputs."""
    import fiksi_b200 as fk
    from fiksi_b200 import workloads as wl
    w = {"truss": lambda: wl.truss(96), "cad_mix": lambda: wl.cad_mix(200), "hinged": lambda: wl.hinged_triangles(8, 64)}[maker]()
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    vg, rg = topo.batch_solve_single_pass(v, p)
    for s in range(0, w.n, 9):
        op, keep = oracle.make_problem(v[s], w.kind, w.idx, p[s], w.free_vars, w.rows)
        vo, ro = oracle.single_pass_problem(op, v[s])
        assert len(ro) == rg.shape[1]
        assert np.array_equal(ro["trace_hash"], rg[s]["trace_hash"]) and np.array_equal(ro["exit_reason"], rg[s]["exit_reason"])
        assert np.max(np.abs(vg[s] - vo)) <= 1e-9 * np.max(np.abs(vo))
    if maker == "hinged":
        assert rg.shape[1] > 1          # the chain really decomposes
