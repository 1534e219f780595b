"""Edge cases of the solve boundary on the GPU: empty batches, problems without free variables or
without rows, a single row, the largest problem of the shared-memory paths and the first one that
takes the global sparse path — each against the CPU oracle."""
import numpy as np
import pytest

import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

pytestmark = pytest.mark.gpu
REL = 1e-9


def _same(oracle, vars_, kind, idx, param, free_vars, rows):
    op, keep = oracle.make_problem(vars_, kind, idx, param, free_vars, rows)
    x0 = np.asarray(vars_, dtype=np.float64)[np.asarray(free_vars, dtype=np.int64)] if len(free_vars) else np.zeros(0)
    xo, ro, trace = oracle.lm_solve(op, x0)
    fp, fkeep = fk.make_problem(vars_, kind, idx, param, free_vars, rows)
    xg, rg = fk.lm_solve(fp, x0)
    assert rg["exit_reason"] == ro["exit_reason"] and rg["trace_hash"] == ro["trace_hash"], (trace, rg, ro)
    assert (rg["outer_iters"], rg["factorizations"], rg["accepted"]) == (ro["outer_iters"], ro["factorizations"], ro["accepted"])
    if len(xo):
        assert np.max(np.abs(xg - xo)) <= REL * max(np.max(np.abs(xo)), 1e-300)
    assert abs(rg["ssr"] - ro["ssr"]) <= REL * max(ro["ssr"], 1.0)
    return rg


def test_empty_batch_is_a_no_op():
    w = wl.truss(4)
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x, rep = topo.batch_solve(v[:0], p[:0])
    assert x.shape == (0, 40) and rep.shape == (0,)
    assert topo.batch_analyze(w.raw_vars[:0], w.raw_param[:0]).shape[0] == 0


def test_single_row_single_free_point(oracle):
    # one distance row, one free point, the other fixed (tests/fixed.rs shape)
    vars_ = np.array([0.0, 0.0, 1.3, 0.4])
    _same(oracle, vars_, [1], [[0, 2, 0, 0]], [1.0], [2, 3], [0])


def test_no_rows_returns_immediately(oracle):
    vars_ = np.array([0.1, 0.2, 1.3, 0.4])
    rg = _same(oracle, vars_, [1], [[0, 2, 0, 0]], [1.0], [0, 1, 2, 3], [])
    assert rg["factorizations"] == 0 and rg["exit_reason"] == 0


def test_no_free_variables(oracle):
    vars_ = np.array([0.0, 0.0, 1.3, 0.4])
    rg = _same(oracle, vars_, [1], [[0, 2, 0, 0]], [1.0], [], [0])
    assert rg["accepted"] == 0


@pytest.mark.parametrize("nx,ny", [(9, 8), (10, 9), (14, 12)])
def test_lattices_around_the_path_boundaries(oracle, nx, ny):
    """Growing lattices cross from the tile path to the CTA path to the global sparse path; every one
    must agree with the oracle whatever path the library picks."""
    w = wl.lattice(nx, ny)
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    _same(oracle, v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    assert topo.info["path"] in (0, 1, 2)


def test_ragged_batch_sizes(oracle):
    w = wl.cad_mix(1)
    for n in (1, 3, 31, 33, 257):
        wn = wl.cad_mix(n)
        v, p, s = wn.prepare()
        topo = fk.Topology.from_arrays(wn.n_vars, wn.kind, wn.idx, wn.free_vars, wn.rows)
        x, rep = topo.batch_solve(v, p)
        op, keep = oracle.make_problem(v[0], wn.kind, wn.idx, p[0], wn.free_vars, wn.rows)
        xo, ro, _ = oracle.lm_solve_batch_uniform(op, v, p, threads=2)
        assert np.array_equal(rep["trace_hash"], ro["trace_hash"]) and np.array_equal(rep["exit_reason"], ro["exit_reason"])
        assert np.max(np.abs(x - xo)) <= REL * np.max(np.abs(xo))


@pytest.mark.parametrize("maker", [lambda: wl.cad_mix(3), lambda: wl.truss(3), lambda: wl.hinged_triangles(4, 3),
                                   lambda: wl.hinged_triangles(16, 3)])
def test_single_system_twin_matches_the_batch_lane_count(maker):
    """fk_topology_lm_solve runs single systems on a 32-lane twin of the topology; the batch entry uses the
    topology's own lane count (4-16 here).  Same operations per entry: coordinates and decisions are identical."""
    w = maker()
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    assert topo.info["path"] == 0 and topo.info["tile"] < 32
    xb, rb = topo.batch_solve(v, p)
    for k in range(len(v)):
        x1, r1 = topo.lm_solve(v[k], p[k], v[k][w.free_vars])
        assert np.array_equal(x1, xb[k])
        assert r1["trace_hash"] == rb[k]["trace_hash"] and r1["exit_reason"] == rb[k]["exit_reason"]
