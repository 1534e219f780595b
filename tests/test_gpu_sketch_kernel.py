"""The sketch-per-thread LM kernel (csrc/lm_sketch.cu: one thread per sketch, state interleaved in shared
memory, tables in the constant bank) against the tile kernel and against the oracle.

The two kernels perform the same LDLt operations in the same order, so traces, exits, counters and lambda are
equal and the coordinates agree to rounding of the back substitution (the sketch kernel gathers, the tile
kernel scatters); against the oracle the usual parity rules of test_gpu_parity.py hold."""
import numpy as np
import pytest

import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

pytestmark = pytest.mark.gpu
REL = 1e-9


def _solve(topo, v, p, kernel):
    with api.lm_kernel(kernel):
        return topo.batch_solve(v, p)


def _topo(w):
    return fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)


@pytest.mark.parametrize("maker,n", [(wl.truss, 4096 + 17), (wl.cad_mix, 4096 + 5), (lambda n: wl.truss(n, n_points=10), 1000),
                                     (lambda n: wl.hinged_triangles(4, n), 333)])
def test_sketch_kernel_equals_tile_kernel_and_oracle(oracle, maker, n):
    w = maker(n)
    v, p, scale = w.prepare()
    topo = _topo(w)
    assert topo.sketch_kernel_info()["available"]
    xs, rs = _solve(topo, v, p, "sketch_solo")
    xt, rt = _solve(topo, v, p, "tile")
    xp, rp = _solve(topo, v, p, "sketch_pair")  # two warps per 32 sketches (falls back to solo when it does not fit)
    for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash", "lambda", "ssr"):
        assert np.array_equal(rs[key], rt[key]), key
        assert np.array_equal(rs[key], rp[key]), key
    assert np.array_equal(xs, xt) and np.array_equal(xs, xp)  # same operations in the same order: bit-identical
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    m = min(n, 1500)
    xo, ro, _ = oracle.lm_solve_batch_uniform(op, v[:m], p[:m], threads=8)
    same = (rs["trace_hash"][:m] == ro["trace_hash"]) & (rs["exit_reason"][:m] == ro["exit_reason"])
    assert same.mean() >= 0.999
    err = np.max(np.abs(xs[:m][same] - xo[same]), axis=1) / np.max(np.abs(xo[same]), axis=1)
    assert err.max() <= REL


def test_sketch_kernel_stress_families(oracle):
    """Config 5 families (singular starts, rank deficiency, 1e20 scales, fixed points, NaN): exits, traces and
    lambda of the sketch kernel equal the tile kernel's, sketch by sketch."""
    for name, w in wl.stress_families(n_each=512):
        v, p, scale = w.prepare(perturb=(name != "nan_coincident_points"))
        topo = _topo(w)
        if not topo.sketch_kernel_info()["available"]:
            continue
        xs, rs = _solve(topo, v, p, "sketch_solo")
        xt, rt = _solve(topo, v, p, "tile")
        xp, rp = _solve(topo, v, p, "sketch_pair")
        for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash"):
            assert np.array_equal(rs[key], rt[key]), (name, key)
            assert np.array_equal(rs[key], rp[key]), (name, key)
        assert np.array_equal(xs, xp, equal_nan=True), name
        assert np.array_equal(rs["lambda"], rt["lambda"], equal_nan=True), name
        fin = np.isfinite(xt).all(axis=1)
        assert np.array_equal(fin, np.isfinite(xs).all(axis=1)), name
        if fin.any():
            d = np.max(np.abs(xs[fin] - xt[fin]), axis=1) / np.maximum(np.max(np.abs(xt[fin]), axis=1), 1e-300)
            assert d.max() <= 1e-9, (name, d.max())


def test_ragged_and_tiny_batches():
    w = wl.truss(70)
    v, p, scale = w.prepare()
    topo = _topo(w)
    xt, rt = _solve(topo, v, p, "tile")
    for n in (1, 31, 32, 33, 70):
        for shape in ("sketch_solo", "sketch_pair"):
            xs, rs = _solve(topo, v[:n], p[:n], shape)
            assert np.array_equal(rs["trace_hash"], rt["trace_hash"][:n])
            assert np.array_equal(xs, xt[:n])


def test_rows_naming_a_variable_twice_stay_on_the_tile_kernel():
    # point-point distance from a point to itself: both slots of x (and of y) are the same free column
    kind = np.array([1, 1], np.uint8)
    idx = np.array([[0, 2, 0, 0], [0, 0, 0, 0]], np.uint32)
    topo = fk.Topology.from_arrays(4, kind, idx, np.arange(4), np.arange(2))
    assert not topo.sketch_kernel_info()["available"]


def _random_topology(rng):
    """A random small system over all 11 expression kinds, some variables fixed (as tests/test_symbolic_parity.py builds them)."""
    n_pts = int(rng.integers(3, 10))
    n_len = int(rng.integers(1, 3))
    n_vars = 2 * n_pts + n_len
    kinds, idxs = [], []
    for _ in range(int(rng.integers(2, 14))):
        k = int(rng.integers(0, 11))
        pts = (2 * rng.choice(n_pts, size=4, replace=n_pts < 4)).tolist()
        if k == 0:
            a, b = rng.choice(n_vars, size=2, replace=False)
            ii = [int(a), int(b), 0, 0]
        elif k == 5:
            ii = pts[:2] + [2 * n_pts + int(rng.integers(0, n_len)), 0]
        elif k == 10:
            ii = pts[:3] + [2 * n_pts + int(rng.integers(0, n_len))]
        else:
            ii = pts
        kinds.append(k)
        idxs.append(ii)
    free = np.sort(rng.choice(n_vars, size=int(rng.integers(max(1, n_vars // 2), n_vars + 1)), replace=False))
    return n_vars, np.array(kinds, np.uint8), np.array(idxs, np.uint32), free.astype(np.uint32), np.arange(len(kinds), dtype=np.uint32)


def test_random_mixed_kind_topologies_all_three_kernels_agree_bit_for_bit(oracle):
    """Forty random systems over all expression kinds with fixed variables, 96 perturbed copies each: tile kernel,
    sketch kernel and its warp-pair shape must agree in every report field and coordinate (NaNs included), and a
    sample must agree with the oracle."""
    rng = np.random.default_rng(77)
    checked = against_oracle = 0
    for trial in range(40):
        n_vars, kind, idx, free, rows = _random_topology(rng)
        topo = fk.Topology.from_arrays(n_vars, kind, idx, free, rows)
        if topo.info["path"] != 0 or not topo.sketch_kernel_info()["available"]:
            continue
        n = 96
        v = rng.normal(size=(1, n_vars)) * 3.0 + rng.normal(size=(n, n_vars)) * 0.05
        v[:, -2:] = np.abs(v[:, -2:]) + 0.5  # lengths / radii positive
        p = np.abs(rng.normal(size=(1, len(kind)))) + 0.3 + np.zeros((n, 1))
        xt, rt = _solve(topo, v, p, "tile")
        xs, rs = _solve(topo, v, p, "sketch_solo")
        xp, rp = _solve(topo, v, p, "sketch_pair")
        for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash"):
            assert np.array_equal(rs[key], rt[key]), (trial, key)
            assert np.array_equal(rp[key], rt[key]), (trial, key)
        assert np.array_equal(rs["lambda"], rt["lambda"], equal_nan=True) and np.array_equal(rs["ssr"], rt["ssr"], equal_nan=True)
        assert np.array_equal(xs, xt, equal_nan=True) and np.array_equal(xp, xt, equal_nan=True), trial
        checked += 1
        if trial % 4 == 0:
            op, keep = oracle.make_problem(v[0], kind, idx, p[0], free, rows)
            xo, ro, _ = oracle.lm_solve_batch_uniform(op, v[:16], p[:16], threads=4)
            same = (rs["trace_hash"][:16] == ro["trace_hash"]) & (rs["exit_reason"][:16] == ro["exit_reason"])
            assert same.mean() >= 0.8, (trial, same.mean())
            ok = same & np.isfinite(xo).all(axis=1) & (ro["exit_reason"] == 0)
            if ok.any():
                err = np.max(np.abs(xs[:16][ok] - xo[ok]), axis=1) / np.maximum(np.max(np.abs(xo[ok]), axis=1), 1e-300)
                assert err.max() <= 1e-6, (trial, err.max())  # (rank-deficient random systems: the converged point is not unique to 1e-9)
            against_oracle += 1
    assert checked >= 25 and against_oracle >= 5
