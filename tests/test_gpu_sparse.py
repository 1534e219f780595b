"""GPU parity of the large single-system path (BASELINE config 3): global-memory evaluation (K1/K2),
normal-equation assembly (K3) and the tree-scheduled sparse LDL^T with sync-free triangular solves
(K5), driven by fk_topology_lm_solve.  Small problems are forced onto this path with
FK_FORCE_PATH=2 so that the CPU oracle can check them."""
import os

import numpy as np
import pytest

import scenarios as sc
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

pytestmark = pytest.mark.gpu
REL = 1e-9


@pytest.fixture
def force_sparse():
    os.environ["FK_FORCE_PATH"] = "2"
    yield
    os.environ.pop("FK_FORCE_PATH", None)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if len(b) else 0.0


def _check_against_oracle(oracle, keep, scale=1.0):
    vars_, kind, idx, param, free_vars, rows = keep
    op, okeep = oracle.make_problem(*keep)
    x0 = vars_[free_vars]
    xo, ro, trace = oracle.lm_solve(op, x0)
    topo = fk.Topology.from_arrays(len(vars_), kind, idx, free_vars, rows)
    assert topo.info["path"] == 2
    xg, rg = topo.lm_solve(vars_, param, x0)
    assert rg["exit_reason"] == ro["exit_reason"], (trace, rg, ro)
    assert rg["trace_hash"] == ro["trace_hash"], (trace, rg, ro)
    assert (rg["outer_iters"], rg["factorizations"], rg["accepted"]) == (ro["outer_iters"], ro["factorizations"], ro["accepted"])
    assert rg["lambda"] == ro["lambda"]
    assert _rel(xg, xo) <= REL, _rel(xg, xo)
    assert abs(rg["ssr"] - ro["ssr"]) <= REL * max(ro["ssr"], 1.0)
    return topo


@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_reference_scenarios_on_sparse_path(oracle, force_sparse, name):
    b = sc.ALL[name](oracle.System)
    for prob, scale, keep in b["s"].prepare(perturb=True):
        _check_against_oracle(oracle, keep)


@pytest.mark.parametrize("nx,ny", [(5, 4), (12, 9)])
def test_small_lattices_forced(oracle, force_sparse, nx, ny):
    w = wl.lattice(nx, ny)
    v, p, scale = w.prepare()
    _check_against_oracle(oracle, (v[0], w.kind, w.idx, p[0], w.free_vars, w.rows))


def test_medium_lattice_natural_path(oracle):
    w = wl.lattice(30, 20)  # 1,200 free variables: takes the sparse path by size
    v, p, scale = w.prepare()
    topo = _check_against_oracle(oracle, (v[0], w.kind, w.idx, p[0], w.free_vars, w.rows))
    t = topo.last_timing()
    assert t["factors"] >= 1 and t["evals"] == t["factors"] + 1


def test_large_eval_bit_exact(oracle):
    w = wl.lattice(40, 30)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    assert topo.info["path"] == 2
    r, j, ms = topo.eval_large(v[0], p[0], v[0][w.free_vars], repeats=3)
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    ro, jo = oracle.evaluate(op, v[0][w.free_vars], jac_nnz=topo.info["jac_nnz"])
    assert np.array_equal(r, ro) and np.array_equal(j, jo) and ms > 0


def test_batch_entry_routes_large_problems(oracle):
    w = wl.lattice(30, 20)
    v, p, scale = w.prepare()
    fp, keep = fk.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    x0 = v[0][w.free_vars]
    xg, rg = fk.lm_solve(fp, x0)
    op, okeep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    xo, ro, _ = oracle.lm_solve(op, x0)
    assert rg["trace_hash"] == ro["trace_hash"] and _rel(xg, xo) <= REL


def test_config3_full_lattice_properties():
    """BASELINE config 3 at full size (200,000 variables, 298,701 rows).  The oracle cannot run it
    (O(m n) scratch per factorisation); size-independent checks: residual exit, every edge at its
    length, a second solve from the solution is a fixed point, deterministic re-run."""
    w = wl.lattice(400, 250)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    i = topo.info
    assert (i["n_free"], i["n_rows"], i["jac_nnz"]) == (200000, 298701, 1194804)
    x0 = v[0][w.free_vars]
    x, rep = topo.lm_solve(v[0], p[0], x0)
    assert rep["exit_reason"] == 0 and rep["ssr"] < 1e-8
    v1 = v[0].copy(); v1[w.free_vars] = x
    r, _, _ = topo.eval_large(v1, p[0], x, want_j=False)
    assert abs(float(np.sum(r * r)) - rep["ssr"]) <= 1e-12 * max(rep["ssr"], 1e-30) + 1e-24
    assert np.max(np.abs(r)) < 1e-4
    x2, rep2 = topo.lm_solve(v1, p[0], x)
    assert rep2["factorizations"] == 0 and np.array_equal(x2, x)
    x3, rep3 = topo.lm_solve(v[0], p[0], x0)
    assert np.array_equal(x3, x) and rep3["trace_hash"] == rep["trace_hash"]


_SCHEDULE_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
import fiksi_b200 as fk
from fiksi_b200 import workloads as wl
w = wl.lattice(150, 120)
v, p, scale = w.prepare()
topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
x, rep = topo.lm_solve(v[0], p[0], v[0][w.free_vars])
np.save(sys.argv[1], x)
print(int(rep["exit_reason"]), int(rep["factorizations"]), int(rep["trace_hash"]), repr(float(rep["ssr"])))
"""


def test_dataflow_and_chained_schedules_match_the_launch_per_step_schedule(tmp_path):
    """The levels near the root are factorised by the tile dataflow kernel and solved by the chained multi-CTA
    kernels, small fronts by one CTA each (csrc/multifrontal.cu).  Against the launch-per-step schedule (FK_NO_FLOW / FK_NO_CHAIN / FK_NO_MID; the knobs are
    read once per process, hence the subprocesses): the dataflow factorisation keeps every summation order, so the
    LM solve is bit-identical; the chained solves use the inverse of the 64x64 pivot triangles, so the coordinates
    agree to rounding."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env in (("default", {}), ("no_flow", {"FK_NO_FLOW": "1"}), ("no_chain", {"FK_NO_CHAIN": "1"}), ("no_mid", {"FK_NO_MID": "1"})):
        path = str(tmp_path / (name + ".npy"))
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", _SCHEDULE_SCRIPT % root, path], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = (r.stdout.strip().splitlines()[-1], np.load(path))
    assert outs["default"][0].split()[0] == "0"
    assert outs["no_flow"][0] == outs["default"][0] and np.array_equal(outs["no_flow"][1], outs["default"][1])
    assert outs["no_chain"][0].split()[:3] == outs["default"][0].split()[:3]
    assert _rel(outs["no_chain"][1], outs["default"][1]) <= 1e-11
    # fronts of at most 72 rows through the 64x64-tile kernels instead of the one-CTA-per-front kernel: other summation
    # orders, same decisions, coordinates to rounding
    assert outs["no_mid"][0].split()[:3] == outs["default"][0].split()[:3]
    assert _rel(outs["no_mid"][1], outs["default"][1]) <= 1e-11


@pytest.mark.parametrize("seed,n_pts,extra", [(1, 150, 40), (2, 400, 150), (3, 700, 0)])
def test_irregular_distance_graphs_on_sparse_path(oracle, force_sparse, seed, n_pts, extra):
    """Irregular structure (not a lattice): random points, every point tied to two earlier nearby points
    (Henneberg steps) plus random extra distances (over-constrained, consistent), a few points pinned; the fronts
    are uneven, so supernodes, dataflow levels and chained solves get shapes the lattices never produce."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 10.0, size=(n_pts, 2))
    edges = [(0, 1)]
    for k in range(2, n_pts):
        d = np.sum((pts[:k] - pts[k]) ** 2, axis=1)
        a, b = np.argsort(d)[:2]
        edges += [(int(a), k), (int(b), k)]
    for _ in range(extra):
        a, b = rng.choice(n_pts, size=2, replace=False)
        edges.append((int(min(a, b)), int(max(a, b))))
    kind = np.ones(len(edges), np.uint8)
    idx = np.array([[2 * a, 2 * b, 0, 0] for a, b in edges], np.uint32)
    dist = np.array([np.hypot(*(pts[a] - pts[b])) for a, b in edges])
    noisy = pts + rng.uniform(-0.02, 0.02, size=pts.shape)
    pinned = {0, 1, 2, 3}  # the first two points stay where they are
    free_vars = np.array([q for q in range(2 * n_pts) if q not in pinned], np.uint32)
    w = wl.Workload("irregular", kind, idx, free_vars, np.arange(len(edges)), noisy.reshape(1, -1), dist.reshape(1, -1))
    v, p, scale = w.prepare()
    _check_against_oracle(oracle, (v[0], w.kind, w.idx, p[0], w.free_vars, w.rows))


def test_large_irregular_graph_properties():
    """24,000 variables of irregular structure at a size the oracle cannot reach: residual exit, every distance met,
    deterministic re-run (the dataflow / chained levels run with uneven fronts)."""
    rng = np.random.default_rng(11)
    n_pts = 12000
    pts = rng.uniform(0.0, 100.0, size=(n_pts, 2))
    order = np.argsort(pts[:, 0] + 0.37 * pts[:, 1], kind="stable")  # sweep order keeps the two neighbours close
    pts = pts[order]
    edges = [(0, 1)]
    for k in range(2, n_pts):
        lo = max(0, k - 400)
        d = np.sum((pts[lo:k] - pts[k]) ** 2, axis=1)
        near = np.argsort(d)[:4] + lo  # four ties per point: over-constrained but consistent, well conditioned
        edges += [(int(a), k) for a in near]
    kind = np.ones(len(edges), np.uint8)
    idx = np.array([[2 * a, 2 * b, 0, 0] for a, b in edges], np.uint32)
    dist = np.array([np.hypot(*(pts[a] - pts[b])) for a, b in edges])
    noisy = pts + rng.uniform(-0.01, 0.01, size=pts.shape)
    free_vars = np.arange(4, 2 * n_pts, dtype=np.uint32)
    w = wl.Workload("irregular_large", kind, idx, free_vars, np.arange(len(edges)), noisy.reshape(1, -1), dist.reshape(1, -1))
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    assert topo.info["path"] == 2
    x0 = v[0][w.free_vars]
    x, rep = topo.lm_solve(v[0], p[0], x0)
    assert rep["exit_reason"] in (0, 2) and rep["ssr"] < 1e-6, rep
    v1 = v[0].copy(); v1[w.free_vars] = x
    r, _, _ = topo.eval_large(v1, p[0], x, want_j=False)
    assert abs(float(np.sum(r * r)) - rep["ssr"]) <= 1e-10 * max(rep["ssr"], 1e-30) + 1e-24
    x2, rep2 = topo.lm_solve(v[0], p[0], x0)
    assert np.array_equal(x2, x) and rep2["trace_hash"] == rep["trace_hash"]
