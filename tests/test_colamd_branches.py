"""COLAMD branches the reference holds no known-answer test for (colamd_rs/src/colamd.rs: dense-column
removal :600-700, garbage collection :1180-1260), product (csrc/colamd_order.hpp) against the oracle's
line-by-line restatement: the permutation, the elimination tree and the R pattern must still be bit-exact.

Every case first proves through the oracle's statistics (stats[1] dense columns, stats[2] garbage
collections) that the branch is really taken (stats[0] counts dense or EMPTIED rows).  Dense ROWS cannot be reached through the solve path: a row of
the augmented Jacobian has at most 8 entries (expression arity), far below max(16, 10 sqrt(n_col))."""
import numpy as np
import pytest

import fiksi_b200 as fk
from test_symbolic_parity import _compare


def _stats(oracle, arrays):
    vars_, kind, idx, param, free_vars, rows = arrays
    op, keep = oracle.make_problem(vars_, kind, idx, param, free_vars, rows)
    sym = oracle.symbolic(op)
    n = len(free_vars)
    ok, p, stats = oracle.colamd(len(rows) + n, n, sym["aug_rowidx"], sym["aug_colptr"])
    assert ok and np.array_equal(p[:n], sym["perm"])
    return stats


def _distance_problem(n_pts, edges, free=None):
    kind = np.ones(len(edges), np.uint8)
    idx = np.array([[2 * a, 2 * b, 0, 0] for a, b in edges], np.uint32)
    free = np.arange(2 * n_pts, dtype=np.uint32) if free is None else np.asarray(free, np.uint32)
    rng = np.random.default_rng(n_pts)
    return rng.normal(size=2 * n_pts), kind, idx, np.ones(len(edges)), free, np.arange(len(edges), dtype=np.uint32)


@pytest.mark.parametrize("n_spokes,n_hubs", [(700, 1), (500, 3), (1500, 2)])
def test_hub_points_are_dense_columns(oracle, n_spokes, n_hubs):
    """A hub point tied to hundreds of others: its two columns have more than max(16, 10 sqrt(min(m, n))) rows
    and are ordered last without taking part in the elimination."""
    n_pts = n_spokes + n_hubs
    edges = [(h, n_hubs + s) for h in range(n_hubs) for s in range(n_spokes)]
    edges += [(n_hubs + s, n_hubs + s + 1) for s in range(n_spokes - 1)]  # a chain through the spokes
    arrays = _distance_problem(n_pts, edges)
    st = _stats(oracle, arrays)
    # stats[0] counts dense AND emptied rows: the hub columns' damping rows lose their only entry (colamd.rs init_scoring)
    assert st[1] == 2 * n_hubs and st[0] == 2 * n_hubs, st[:4]
    topo = _compare(oracle, arrays)
    perm = topo.symbolic()["perm"]
    assert sorted(perm[-2 * n_hubs:].tolist()) == list(range(2 * n_hubs))  # the hub columns close the ordering


def test_hub_with_fixed_coordinates(oracle):
    """Half of the hub is fixed: one dense column instead of two, and rows that lose entries."""
    n_spokes = 800
    edges = [(0, 1 + s) for s in range(n_spokes)] + [(1 + s, 2 + s) for s in range(n_spokes - 1)]
    free = [q for q in range(2 * (n_spokes + 1)) if q != 1]
    arrays = _distance_problem(n_spokes + 1, edges, free)
    assert _stats(oracle, arrays)[1] == 1
    _compare(oracle, arrays)


def _random_graph(rng, n_pts, n_edges):
    edges = set()
    while len(edges) < n_edges:
        a, b = (int(v) for v in rng.integers(0, n_pts, size=2))
        if a != b:
            edges.add((min(a, b), max(a, b)))
    return sorted(edges)


def test_garbage_collection_is_exercised(oracle):
    """Dense-ish random graphs fill the workspace colamd_recommended() grants (the size the LM path passes,
    solvi qr.rs:140-150) and force the reference to compact it; the product orders columns without that
    workspace but must produce the same permutation."""
    rng = np.random.default_rng(2024)
    hits = 0
    for trial in range(40):
        n_pts = int(rng.integers(30, 90))
        n_edges = int(n_pts * rng.uniform(3.0, 7.0))
        arrays = _distance_problem(n_pts, _random_graph(rng, n_pts, min(n_edges, n_pts * (n_pts - 1) // 2)))
        if _stats(oracle, arrays)[2] > 0:
            hits += 1
            _compare(oracle, arrays)
            if hits >= 6:
                break
    assert hits >= 3, "no trial made the reference collect garbage: the test lost its teeth"
