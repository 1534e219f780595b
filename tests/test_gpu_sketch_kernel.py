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
