"""Host-side supernodal multifrontal analysis (fiksi_b200/csrc/multifrontal.cu::build_symbolic) checked
on the CPU against the R pattern the oracle derives from solvi's CholeskyStructure
(solvi/src/decomposition/sparse/cholesky.rs:359-595): supernode partition, front row lists,
relative indices, small-subtree / level classification and the static tile task lists."""
import numpy as np
import pytest

import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

TB = 64


def _l_columns(sym, n):
    """Columns of L (= rows of R) as sorted lists, from the R pattern (per column ascending rows)."""
    cols = [[] for _ in range(n)]
    rc, ri = sym["r_colptr"], sym["r_rowidx"]
    for j in range(n):
        for q in range(rc[j], rc[j + 1]):
            cols[ri[q]].append(j)      # R(k, j) != 0  <=>  L(j, k) != 0
    return cols


@pytest.mark.parametrize("nx,ny", [(6, 5), (30, 20), (60, 45)])
def test_supernodes_fronts_and_relative_indices(nx, ny):
    w = wl.lattice(nx, ny)
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    n = topo.info["n_free"]
    sym = topo.symbolic()
    sn = topo.supernodal()
    S = sn["info"]["n_supernodes"]
    first, front, par = sn["sn_first"], sn["front"], sn["sn_parent"]
    assert first[0] == 0 and first[S] == n and np.all(np.diff(first.astype(np.int64)) > 0)
    cols = _l_columns(sym, n)
    etree = sym["etree_parent"]
    rows_off = np.concatenate([[0], np.cumsum(front)]).astype(np.int64)
    assert rows_off[-1] == sn["info"]["rows_total"]
    col2sn = np.repeat(np.arange(S), np.diff(first.astype(np.int64)))
    rel_off = 0
    for s in range(S):
        c0, c1 = int(first[s]), int(first[s + 1])
        rows = sn["rows"][rows_off[s]:rows_off[s + 1]].tolist()
        assert rows == cols[c0], s                                  # the front is the structure of the first column
        for j in range(c0 + 1, c1):                                 # nested structures inside a supernode
            assert cols[j] == rows[j - c0:]
            assert etree[j - 1] == j
        last = c1 - 1
        expect_parent = -1 if etree[last] < 0 else int(col2sn[etree[last]])
        assert par[s] == expect_parent
        r = len(rows) - (c1 - c0)
        if par[s] >= 0:
            prow = sn["rows"][rows_off[par[s]]:rows_off[par[s] + 1]]
            rel = sn["rel"][rel_off:rel_off + r]
            assert np.array_equal(prow[rel], np.array(rows[c1 - c0:], dtype=np.uint32))
            assert np.all(np.diff(rel.astype(np.int64)) > 0)
        else:
            assert r == 0
        rel_off += r
    assert rel_off == sn["info"]["rel_total"]
    assert sn["info"]["panel_doubles"] == int(np.sum(front.astype(np.int64) * np.diff(first.astype(np.int64))))
    # classification: fronts above 32 are big, a parent of a big supernode is big, levels grow towards the root
    big, level = sn["big"], sn["level"]
    assert np.all(big[front > 32] == 1)
    for s in range(S):
        if big[s] and par[s] >= 0:
            assert big[par[s]] and level[par[s]] > level[s]
        if not big[s]:
            assert front[s] <= 32


def test_task_lists_cover_every_tile_once():
    w = wl.lattice(60, 45)
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    sn = topo.supernodal()
    first, front, big, level = sn["sn_first"], sn["front"], sn["big"], sn["level"]
    width = np.diff(first.astype(np.int64))
    tasks, launches = sn["tasks"], sn["launches"]
    assert sn["info"]["n_big"] > 0 and len(launches) > 0
    seen_diag, seen_col, seen_upd = set(), set(), {}
    prev_level = 0
    for kind, t0, cnt in launches:
        for s, a, b, pk in tasks[t0:t0 + cnt]:
            assert big[s]
            assert level[s] >= prev_level                      # launches go level by level
            f, ns = int(front[s]), int(width[s])
            if kind == 1:                                      # diagonal tile of pivot block b
                assert a == b and b % TB == 0 and b < ns and (pk >> 16) == min(TB, ns - b)
                assert (s, b) not in seen_diag
                seen_diag.add((s, b))
            elif kind == 2:                                    # panel tile rows a.., pivot block b
                assert (s, b) in seen_diag and a >= b + (pk >> 16) and a + (pk & 0xFFFF) <= f
                assert (s, a, b) not in seen_col
                seen_col.add((s, a, b))
            elif kind == 3:                                    # right-looking update of tile (a, b) by pivot block kb
                kb = pk >> 24
                nra, nrb, k = (pk & 0xFF) + 1, ((pk >> 8) & 0xFF) + 1, ((pk >> 16) & 0xFF) + 1
                assert k == min(TB, ns - kb * TB) and a >= b >= kb * TB + k and a + nra <= f and b + nrb <= f
                assert (s, a, kb * TB) in seen_col and ((s, b, kb * TB) in seen_col)
                seen_upd.setdefault((s, a, b), []).append(kb)
            prev_level = max(prev_level, int(level[s]))
    # every big supernode: all pivot blocks factorised, every tile right of / below block kb updated by kb exactly once
    for s in np.flatnonzero(big):
        f, ns = int(front[s]), int(width[s])
        nsb = (ns + TB - 1) // TB
        assert {b for (ss, b) in seen_diag if ss == s} == {kb * TB for kb in range(nsb)}
        starts = [kb * TB for kb in range(nsb)] + list(range(ns, f, TB))
        for bi, a in enumerate(starts):
            for b in starts[:bi + 1]:
                want = [kb for kb in range(nsb) if kb * TB < b] if b < ns else list(range(nsb))
                want = [kb for kb in want if kb * TB + min(TB, ns - kb * TB) <= b]
                assert seen_upd.get((s, a, b), []) == want, (s, a, b)


def test_tiny_system_has_only_small_subtrees():
    w = wl.lattice(3, 3)
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    sn = topo.supernodal()
    assert sn["info"]["n_big"] == 0 and sn["info"]["n_launches"] == 0 and sn["info"]["n_small_subtrees"] >= 1
    assert sn["sn_first"][-1] == topo.info["n_free"]
