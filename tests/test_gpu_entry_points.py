"""The drop-in entry points (fk_lm_solve, fk_lm_solve_batch) as a caller that owns flattened problems uses them:
symbolic analysis once per topology through the process-wide cache, heterogeneous batches with all topologies in
flight together, pooled buffers on the L-BFGS / analyze entries, concurrent callers."""
import threading


import numpy as np
import pytest

import fiksi_b200 as fk
from fiksi_b200 import api, workloads as wl

pytestmark = pytest.mark.gpu
REL = 1e-9


def _problems(w, v, p, count):
    return [fk.make_problem(v[k], w.kind, w.idx, p[k], w.free_vars, w.rows) for k in range(count)]


def test_lm_solve_reuses_the_symbolic_analysis(oracle):
    api.topology_cache_clear()
    w = wl.hinged_triangles(16, 3)
    v, p, scale = w.prepare()
    probs = _problems(w, v, p, 3)
    h0, m0, _ = api.topology_cache_stats()
    x_cold, r_cold = fk.lm_solve(probs[0][0], v[0][w.free_vars])
    h1, m1, e1 = api.topology_cache_stats()
    assert (h1 - h0, m1 - m0) == (0, 1) and e1 >= 1
    x_warm, r_warm = fk.lm_solve(probs[0][0], v[0][w.free_vars])
    h2, m2, _ = api.topology_cache_stats()
    assert (h2 - h1, m2 - m1) == (1, 0)
    assert np.array_equal(x_cold, x_warm) and r_cold == r_warm
    # a structurally different problem (one more fixed coordinate) is a different topology
    fp, keep = fk.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars[1:], w.rows)
    fk.lm_solve(fp, v[0][w.free_vars[1:]])
    h3, m3, _ = api.topology_cache_stats()
    assert (h3 - h2, m3 - m2) == (0, 1)
    op, okeep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    xo, ro, _ = oracle.lm_solve(op, v[0][w.free_vars])
    assert r_warm["trace_hash"] == ro["trace_hash"] and np.max(np.abs(x_warm - xo)) <= REL * np.max(np.abs(xo))


def test_cache_can_be_disabled_and_bounded():
    api.topology_cache_clear()
    api.topology_cache_configure(0)
    try:
        w = wl.truss(1, n_points=6)
        v, p, scale = w.prepare()
        fp, keep = fk.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
        for _ in range(2):
            fk.lm_solve(fp, v[0][w.free_vars])
        assert api.topology_cache_stats()[2] == 0
        api.topology_cache_configure(2)
        for n_points in (5, 6, 7, 8):
            w = wl.truss(1, n_points=n_points)
            v, p, scale = w.prepare()
            fp, keep = fk.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
            fk.lm_solve(fp, v[0][w.free_vars])
        assert api.topology_cache_stats()[2] == 2
    finally:
        api.topology_cache_configure(1024)


def test_heterogeneous_batch_matches_the_oracle_problem_by_problem(oracle):
    """Some thirty different small topologies with a few members each, interleaved in one fk_lm_solve_batch call."""
    api.topology_cache_clear()
    items = []
    for n_points in range(4, 24):
        for maker in (lambda c, n=n_points: wl.truss(c, n_points=n), lambda c, n=n_points: wl.hinged_triangles(max(1, n // 3), c)):
            w = maker(3)
            v, p, scale = w.prepare()
            v = v * (1.0 + 0.01 * np.arange(3))[:, None]  # three different starting points per topology
            for k in range(3):
                items.append((w, v[k], p[k]))
    order = np.random.default_rng(0).permutation(len(items))
    probs, keeps, x0s = [], [], []
    for i in order:
        w, vk, pk = items[i]
        fp, keep = fk.make_problem(vk, w.kind, w.idx, pk, w.free_vars, w.rows)
        probs.append(fp); keeps.append(keep); x0s.append(vk[w.free_vars])
    xs, reps = fk.lm_solve_batch(probs, x0s)
    _, misses, entries = api.topology_cache_stats()
    distinct = len({w.name for w, _, _ in items})
    assert entries == distinct and distinct >= 25
    for j, i in enumerate(order):
        w, vk, pk = items[i]
        op, okeep = oracle.make_problem(vk, w.kind, w.idx, pk, w.free_vars, w.rows)
        xo, ro, _ = oracle.lm_solve(op, vk[w.free_vars])
        assert reps["trace_hash"][j] == ro["trace_hash"] and reps["exit_reason"][j] == ro["exit_reason"], (j, w.name)
        assert np.max(np.abs(xs[j] - xo)) <= REL * max(np.max(np.abs(xo)), 1e-300)
    # the same batch again: no analysis at all
    h0, m0, _ = api.topology_cache_stats()
    xs2, reps2 = fk.lm_solve_batch(probs, x0s)
    h1, m1, _ = api.topology_cache_stats()
    assert m1 == m0 and h1 - h0 == distinct
    assert all(np.array_equal(a, b) for a, b in zip(xs, xs2))


def test_null_arrays_are_rejected_not_dereferenced():
    w = wl.truss(1, n_points=5)
    v, p, scale = w.prepare()
    fp, keep = fk.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    fp.vars = None
    with pytest.raises(fk.FiksiError) as e:
        fk.lm_solve(fp, v[0][w.free_vars])
    assert e.value.code == -1


def test_repeated_lbfgs_and_analyze_calls_reuse_their_buffers(oracle):
    w = wl.truss(600)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x1, r1 = topo.batch_solve_lbfgs(v, p)
    x2, r2 = topo.batch_solve_lbfgs(v[:100], p[:100])       # smaller, then larger again
    x3, r3 = topo.batch_solve_lbfgs(v, p)
    assert np.array_equal(x1, x3) and np.array_equal(x1[:100], x2) and np.array_equal(r1["trace_hash"], r3["trace_hash"])
    a1 = topo.batch_analyze(w.raw_vars[:50], w.raw_param[:50])
    a2 = topo.batch_analyze(w.raw_vars, w.raw_param)
    a3 = topo.batch_analyze(w.raw_vars[:50], w.raw_param[:50])
    assert np.array_equal(a1, a3) and np.array_equal(a1, a2[:50])


def test_concurrent_callers_on_one_large_topology():
    """Two host threads solve different starting points on the same path-2 topology: the solver's work vectors and
    CUDA graphs are shared, so the calls must serialise inside the library and return what sequential calls return."""
    w = wl.lattice(40, 30)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    if topo.info["path"] != 2:
        pytest.skip("needs the global sparse path")
    starts = [v[0], v[0] * 1.001, v[0] * 0.999, v[0] * 1.002]
    ref = [topo.lm_solve(s, p[0], s[w.free_vars]) for s in starts]
    out = [None] * len(starts)

    def work(k):
        out[k] = topo.lm_solve(starts[k], p[0], starts[k][w.free_vars])
    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(starts))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for (x, r), (xr, rr) in zip(out, ref):
        assert np.array_equal(x, xr) and r == rr


@pytest.mark.parametrize("maker,shared", [(wl.truss, True), (wl.truss, False), (wl.cad_mix, False),
                                          (lambda n: wl.hinged_triangles(4, n), True)])
def test_batch_system_solve_does_the_pre_and_post_processing_on_the_device(oracle, maker, shared):
    """fk_batch_system_solve == Workload.prepare (the bit-exact numpy restatement of assemble::solve's scale and seeded
    perturbation, itself pinned to the oracle by test_workload_prepare_matches_assemble) + fk_batch_solve + write-back:
    same scales bit for bit, same traces, same unscaled coordinates bit for bit."""
    n = 5000
    w = maker(n)
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x_ref, rep_ref = topo.batch_solve(v, p)
    want = w.write_back(w.raw_vars, x_ref, scale)[:, w.free_vars]
    raw_param = w.raw_param[0] if shared else w.raw_param
    if shared:
        assert np.all(w.raw_param == w.raw_param[0])
    x, scales, rep = topo.batch_system_solve(w.raw_vars, raw_param, perturb=True, shared_param=shared)
    assert np.array_equal(scales, scale)
    for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash", "lambda", "ssr"):
        assert np.array_equal(rep[key], rep_ref[key]), key
    assert np.array_equal(x, want)
    # without the perturbation
    v0, p0, s0 = w.prepare(perturb=False)
    x0_ref, rep0_ref = topo.batch_solve(v0, p0)
    x0, sc0, rep0 = topo.batch_system_solve(w.raw_vars, raw_param, perturb=False, shared_param=shared)
    assert np.array_equal(rep0["trace_hash"], rep0_ref["trace_hash"])
    assert np.array_equal(x0, w.write_back(w.raw_vars, x0_ref, s0)[:, w.free_vars])


def test_batch_system_solve_in_two_halves_keeps_two_batches_in_flight():
    """fk_batch_system_solve_begin / _wait: batches A, B, C, D streamed two deep (different inputs, different parameter
    rows, own output buffers) return exactly what the one-call form returns for each of them."""
    import torch
    n = 20000
    topo = None
    batches = []
    for k in range(4):
        w = wl.truss(n, first=k * n)
        if topo is None:
            topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
        raw = np.ascontiguousarray(w.raw_vars * (1.0 + 0.25 * k))  # (another scale per batch: other values, same topology)
        par = np.ascontiguousarray(w.raw_param[0] * (1.0 + 0.25 * k))
        x_ref, sc_ref, rep_ref = topo.batch_system_solve(raw, par, perturb=True, shared_param=True)
        nf = topo.info["n_free"]
        bufs = (torch.from_numpy(raw).pin_memory(), torch.from_numpy(par).pin_memory(),
                torch.zeros((n, nf), dtype=torch.float64).pin_memory(), torch.zeros((n, 5), dtype=torch.float64).pin_memory())
        batches.append((bufs, x_ref, rep_ref))
    tokens = []
    for k, (bufs, _, _) in enumerate(batches):
        tokens.append(topo.batch_system_solve_begin(0, n, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
                                                    shared_param=True))
        if k >= 1:
            topo.batch_system_solve_wait(tokens[k - 1])
    topo.batch_system_solve_wait(tokens[-1])
    topo.batch_system_solve_wait(tokens[0])  # waiting again (or for an old token) is harmless
    for bufs, x_ref, rep_ref in batches:
        rep = bufs[3].numpy().view(fk.REPORT_DTYPE).reshape(n)
        assert np.array_equal(bufs[2].numpy(), x_ref)
        for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash", "lambda", "ssr"):
            assert np.array_equal(rep[key], rep_ref[key]), key
    with pytest.raises(fk.FiksiError):
        topo.batch_system_solve_wait(tokens[-1] + 1000)


def test_batch_solve_device_in_two_halves():
    """fk_batch_solve_device_begin / _wait on three different batches of the mixed CAD topology, two in flight."""
    import torch
    n = 30001
    batches = []
    topo = None
    for k in range(3):
        w = wl.cad_mix(n, first=k * n)
        v, p, scale = w.prepare()
        if topo is None:
            topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
        x_ref, rep_ref = topo.batch_solve(v, p)
        bufs = (torch.from_numpy(v).pin_memory(), torch.from_numpy(p).pin_memory(),
                torch.zeros((n, topo.info["n_free"]), dtype=torch.float64).pin_memory(), torch.zeros((n, 5), dtype=torch.float64).pin_memory())
        batches.append((bufs, x_ref, rep_ref))
    tokens = []
    for k, (bufs, _, _) in enumerate(batches):
        tokens.append(topo.batch_solve_begin(0, n, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr()))
        if k >= 1:
            topo.batch_solve_wait(tokens[k - 1])
    topo.batch_solve_wait(tokens[-1])
    for bufs, x_ref, rep_ref in batches:
        rep = bufs[3].numpy().view(fk.REPORT_DTYPE).reshape(n)
        assert np.array_equal(bufs[2].numpy(), x_ref)
        for key in ("exit_reason", "outer_iters", "factorizations", "accepted", "trace_hash", "lambda", "ssr"):
            assert np.array_equal(rep[key], rep_ref[key]), key


def test_batch_system_solve_with_fixed_variables_and_a_custom_perturbation_list():
    name, w = [f for f in wl.stress_families(n_each=700) if f[0] == "two_components_a"][0]
    v, p, scale = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x_ref, rep_ref = topo.batch_solve(v, p)
    x, scales, rep = topo.batch_system_solve(w.raw_vars, w.raw_param, perturb=True, perturb_vars=w.perturb_vars)
    assert np.array_equal(scales, scale) and np.array_equal(rep["trace_hash"], rep_ref["trace_hash"])
    assert np.array_equal(x, scale[:, None] * x_ref)
