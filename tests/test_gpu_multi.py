"""Sketch sharding across the devices of one node: contiguous ranges per device, no
data-path collective, results identical to a single-device run.  Needs >= 2 visible GPUs; skipped
otherwise (the range bookkeeping itself is covered on CPU by tests/test_sharding_gloo.py)."""
import numpy as np
import pytest

import fiksi_b200 as fk
from fiksi_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _need_two():
    if fk.device_count() < 2:
        pytest.skip("needs two visible GPUs")


def test_uniform_batch_sharded_over_two_devices():
    _need_two()
    for maker, n in ((wl.truss, 4099), (wl.cad_mix, 10001)):
        w = maker(n)
        v, p, s = w.prepare()
        topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
        x1, r1 = topo.batch_solve(v, p, n_gpus=1)
        x2, r2 = topo.batch_solve(v, p, n_gpus=2)
        assert np.array_equal(x1, x2)
        for key in r1.dtype.names:
            assert np.array_equal(r1[key], r2[key]), key


def test_mixed_problem_list_sharded_over_two_devices(oracle):
    _need_two()
    probs, x0s, keeps = [], [], []
    for k, maker in enumerate((wl.truss, wl.cad_mix, wl.truss, wl.cad_mix, wl.cad_mix)):
        w = maker(3, first=10 * k) if maker is wl.truss else maker(3)
        v, p, s = w.prepare()
        for j in range(3):
            fp, keep = fk.make_problem(v[j], w.kind, w.idx, p[j], w.free_vars, w.rows)
            probs.append(fp); keeps.append(keep); x0s.append(v[j][w.free_vars])
    xa, ra = fk.lm_solve_batch(probs, x0s, n_gpus=1)
    xb, rb = fk.lm_solve_batch(probs, x0s, n_gpus=2)
    for a, b in zip(xa, xb):
        assert np.array_equal(a, b)
    assert np.array_equal(ra["trace_hash"], rb["trace_hash"])


def test_second_device_explicitly():
    _need_two()
    w = wl.truss(513)
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    x0, r0 = topo.batch_solve(v, p)
    out = np.zeros_like(x0); rep = np.zeros(len(x0), dtype=fk.REPORT_DTYPE)
    topo.batch_solve_into(1, len(x0), v.ctypes.data, p.ctypes.data, out.ctypes.data, rep.ctypes.data)
    assert np.array_equal(out, x0) and np.array_equal(rep["trace_hash"], r0["trace_hash"])
    assert np.array_equal(topo.batch_analyze(w.raw_vars, w.raw_param, device=1), topo.batch_analyze(w.raw_vars, w.raw_param, device=0))
    xl0, rl0 = topo.batch_solve_lbfgs(v, p, device=0)
    xl1, rl1 = topo.batch_solve_lbfgs(v, p, device=1)
    assert np.array_equal(xl0, xl1)


def test_many_topologies_go_to_devices_as_whole_groups(oracle):
    """More topologies than devices: fk_lm_solve_batch hands whole groups to the devices round robin; every problem
    must still get the answer a single-device call gives."""
    _need_two()
    probs, x0s, keeps = [], [], []
    for n_points in range(4, 16):
        w = wl.truss(2, n_points=n_points)
        v, p, s = w.prepare()
        for j in range(2):
            fp, keep = fk.make_problem(v[j], w.kind, w.idx, p[j], w.free_vars, w.rows)
            probs.append(fp); keeps.append(keep); x0s.append(v[j][w.free_vars])
    xa, ra = fk.lm_solve_batch(probs, x0s, n_gpus=1)
    xb, rb = fk.lm_solve_batch(probs, x0s, n_gpus=2)
    for a, b in zip(xa, xb):
        assert np.array_equal(a, b)
    assert np.array_equal(ra["trace_hash"], rb["trace_hash"]) and np.array_equal(ra["exit_reason"], rb["exit_reason"])


def test_single_pass_and_system_solve_on_two_devices():
    _need_two()
    w = wl.hinged_triangles(8, 3000)
    v, p, s = w.prepare()
    topo = fk.Topology.from_arrays(w.n_vars, w.kind, w.idx, w.free_vars, w.rows)
    v1, r1 = topo.batch_solve_single_pass(v, p, n_gpus=1)
    v2, r2 = topo.batch_solve_single_pass(v, p, n_gpus=2)
    assert np.array_equal(v1, v2) and np.array_equal(r1["trace_hash"], r2["trace_hash"])
    xa, sa, ra = topo.batch_system_solve(w.raw_vars, w.raw_param, device=0)
    xb, sb, rb = topo.batch_system_solve(w.raw_vars, w.raw_param, device=1)
    assert np.array_equal(xa, xb) and np.array_equal(sa, sb) and np.array_equal(ra["trace_hash"], rb["trace_hash"])


def test_heterogeneous_batch_split_over_two_devices():
    """Enough different small systems for both devices: each device gets a contiguous share and one launch of the
    heterogeneous kernel; the answers are those of a single-device call."""
    _need_two()
    probs, x0s, keeps = [], [], []
    for n_points in range(4, 20):
        w = wl.truss(10, n_points=n_points)
        v, p, s = w.prepare()
        for j in range(10):
            fp, keep = fk.make_problem(v[j], w.kind, w.idx, p[j], w.free_vars, w.rows)
            probs.append(fp); keeps.append(keep); x0s.append(v[j][w.free_vars])
    xa, ra = fk.lm_solve_batch(probs, x0s, n_gpus=1)
    xb, rb = fk.lm_solve_batch(probs, x0s, n_gpus=2)
    for a, b in zip(xa, xb):
        assert np.array_equal(a, b)
    assert np.array_equal(ra["trace_hash"], rb["trace_hash"])
