"""Product host pipeline vs oracle, no GPU needed: the augmented-Jacobian CSC pattern, the COLAMD
permutation, the column elimination tree and the R pattern must be bit-exact (north_star parity
objects #1 and #2), on every reference scenario, the BASELINE configs' topologies
and random sparse problems."""
import numpy as np
import pytest

import scenarios as sc
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl


def _compare(oracle, prob_arrays):
    vars_, kind, idx, param, free_vars, rows = prob_arrays
    op, keep = oracle.make_problem(vars_, kind, idx, param, free_vars, rows)
    ref = oracle.symbolic(op)
    topo = fk.Topology.from_arrays(len(vars_), kind, idx, free_vars, rows)
    got = topo.symbolic()
    for key in ("aug_colptr", "aug_rowidx", "perm", "etree_parent", "r_colptr", "r_rowidx"):
        assert np.array_equal(ref[key], got[key]), key
    assert sorted(got["perm"].tolist()) == list(range(len(free_vars)))
    return topo


@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_reference_scenarios(oracle, name):
    b = sc.ALL[name](oracle.System)
    for prob, scale, keep in b["s"].prepare(perturb=True):
        _compare(oracle, keep)


def test_truss_topology(oracle):
    w = wl.truss(4)
    v, p, scale = w.prepare()
    topo = _compare(oracle, (v[0], w.kind, w.idx, p[0], w.free_vars, w.rows))
    i = topo.info
    assert (i["n_free"], i["n_rows"], i["jac_nnz"], i["aug_nnz"]) == (40, 37, 148, 188)
    assert i["path"] == 0 and i["tile"] == 16


def test_cad_mix_topology_matches_reference_example(oracle):
    """The flattened circle_triangle_line arrays of workloads.cad_mix are what fiksi's own
    constructors produce (examples/fiksi_svg_tests/src/main.rs:12-45)."""
    b = sc.circle_triangle_line(oracle.System)
    (prob, scale, keep), = b["s"].prepare(perturb=False)
    w = wl.cad_mix(3)
    assert np.array_equal(keep[1], w.kind) and np.array_equal(keep[2], w.idx)
    assert np.array_equal(keep[4], w.free_vars) and np.array_equal(keep[5], w.rows)
    topo = _compare(oracle, (keep[0], w.kind, w.idx, keep[3], w.free_vars, w.rows))
    i = topo.info
    assert (i["n_free"], i["n_rows"], i["jac_nnz"]) == (11, 8, 45)


def test_workload_prepare_matches_assemble(oracle):
    """Workload.prepare (numpy, vectorised) == assemble::solve's scale + perturbation, bit for bit."""
    b = sc.circle_triangle_line(oracle.System)
    (prob, scale, keep), = b["s"].prepare(perturb=True)
    w = wl.Workload("ctl", wl._CTL_KIND, wl._CTL_IDX, np.arange(11), np.arange(8), wl._CTL_BASE.reshape(1, -1),
                    wl._CTL_PARAM.reshape(1, -1))
    v, p, s = w.prepare(perturb=True)
    assert s[0] == scale
    assert np.array_equal(v[0], keep[0]) and np.array_equal(p[0], keep[3])
    # hinged triangles of the reference's criterion bench
    for n in (1, 4):
        b = sc.hinged_triangles_bench(oracle.System, n)
        (prob, scale, keep), = b["s"].prepare(perturb=True)
        w = wl.hinged_triangles(n)
        v, p, s = w.prepare(perturb=True)
        assert s[0] == scale and np.array_equal(v[0], keep[0]) and np.array_equal(p[0], keep[3])
        assert np.array_equal(keep[1], w.kind) and np.array_equal(keep[2], w.idx)


def test_lattice_small_and_medium(oracle):
    for nx, ny in ((5, 4), (30, 20)):
        w = wl.lattice(nx, ny)
        v, p, scale = w.prepare()
        topo = _compare(oracle, (v[0], w.kind, w.idx, p[0], w.free_vars, w.rows))
        assert topo.info["n_free"] == 2 * nx * ny


def test_random_problems(oracle):
    rng = np.random.default_rng(5)
    for trial in range(60):
        n_pts = int(rng.integers(2, 14))
        n_len = int(rng.integers(0, 3))
        n_vars = 2 * n_pts + n_len
        kinds, idxs = [], []
        for _ in range(int(rng.integers(1, 25))):
            k = int(rng.integers(0, 11))
            pts = (2 * rng.integers(0, n_pts, size=4)).tolist()
            if k == 0:
                ii = [int(rng.integers(0, n_vars)), int(rng.integers(0, n_vars)), 0, 0]
            elif k in (5, 10):
                if n_len == 0:
                    continue
                ii = pts[:2] + [2 * n_pts + int(rng.integers(0, n_len))] + [0] if k == 5 else pts[:3] + [2 * n_pts + int(rng.integers(0, n_len))]
            else:
                ii = pts
            kinds.append(k)
            idxs.append(ii)
        if not kinds:
            continue
        free = np.sort(rng.choice(n_vars, size=int(rng.integers(1, n_vars + 1)), replace=False))
        rows = np.sort(rng.choice(len(kinds), size=int(rng.integers(1, len(kinds) + 1)), replace=False))
        vars_ = rng.normal(size=n_vars)
        _compare(oracle, (vars_, kinds, idxs, rng.normal(size=len(kinds)), free, rows))


def test_invalid_problems_are_rejected():
    with pytest.raises(fk.FiksiError):
        fk.Topology.from_arrays(4, [13], [[0, 2, 0, 0]], [0, 1], [0])      # unknown kind (11 and 12 are the pose rows of ClusteredSystem)
    with pytest.raises(fk.FiksiError):
        fk.Topology.from_arrays(4, [1], [[0, 4, 0, 0]], [0, 1], [0])       # variable out of range
    with pytest.raises(fk.FiksiError):
        fk.Topology.from_arrays(4, [1], [[0, 2, 0, 0]], [0, 0], [0])       # duplicate free variable
    with pytest.raises(fk.FiksiError):
        fk.Topology.from_arrays(4, [1], [[0, 2, 0, 0]], [0, 1], [3])       # row out of range


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(fk.LIB_PATH), "..", "include", "fiksi_b200.h")).read()
    declared = set(re.findall(r"FK_API\s+[^;(]*?\b(fk_\w+)\s*\(", hdr))
    assert declared and declared == set(fk._lib.EXPORTS)
    L = fk.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_threaded_symbolic_matches_single_thread_and_oracle(oracle, monkeypatch):
    """Large systems run the R pattern, the L transposition and the H lookups on host threads
    (csrc/symbolic.cpp); the arrays must not depend on the thread count, and must still be the
    reference's (cholesky.rs:359-595 via the oracle)."""
    w = wl.lattice(120, 100)  # 24,000 variables: every threaded phase is active
    v, p, scale = w.prepare()
    outs = []
    for threads in ("1", "7"):
        monkeypatch.setenv("FK_SYM_THREADS", threads)
        topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
        assert topo.info["path"] == 2
        outs.append((topo.symbolic(), topo.supernodal()))
    (s1, n1), (s7, n7) = outs
    for key in ("aug_colptr", "aug_rowidx", "perm", "etree_parent", "r_colptr", "r_rowidx"):
        assert np.array_equal(s1[key], s7[key]), key
    for key in ("sn_first", "front", "sn_parent", "rows", "rel", "big", "level", "tasks", "launches"):
        assert np.array_equal(n1[key], n7[key]), key
    ref = oracle.symbolic(oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)[0])
    assert np.array_equal(ref["perm"], s7["perm"])
    assert np.array_equal(ref["r_colptr"], s7["r_colptr"]) and np.array_equal(ref["r_rowidx"], s7["r_rowidx"])
