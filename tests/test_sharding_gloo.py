"""N > 1 path on CPU: two gloo ranks shard a batch by sketch, solve their ranges independently
and gather; the concatenation must equal the single-rank result bit for bit (sketches are
independent).  The per-shard solver here is the CPU oracle (test infrastructure) — on
GPUs it is Topology.batch_solve."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from fiksi_b200 import workloads as wl
    from fiksi_b200.distributed import shard_bounds, solve_sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w = wl.truss(n)
    v, p, scale = w.prepare()
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)

    def solver(vs, ps):
        x, rep, _ = oracle.lm_solve_batch_uniform(op, vs, ps, threads=2)
        return x, rep

    free, reports = solve_sharded(solver, v, p, n_free=40)
    assert free.shape == (n, 40) and len(reports) == n
    lo, hi = shard_bounds(n, world)[rank]
    local, _ = solve_sharded(solver, v, p, n_free=40, gather=False)
    assert np.array_equal(local, free[lo:hi])
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "free.npy"), free)
        np.save(os.path.join(out_dir, "trace.npy"), reports["trace_hash"])
    dist.destroy_process_group()


def test_two_ranks_concatenate_to_single_rank(tmp_path, oracle):
    from fiksi_b200 import workloads as wl
    from fiksi_b200.distributed import shard_bounds

    n = 301  # odd on purpose: ragged shards
    assert shard_bounds(n, 2) == [(0, 150), (150, 301)] and shard_bounds(5, 8)[-1] == (4, 5)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    w = wl.truss(n)
    v, p, scale = w.prepare()
    op, keep = oracle.make_problem(v[0], w.kind, w.idx, p[0], w.free_vars, w.rows)
    x, rep, _ = oracle.lm_solve_batch_uniform(op, v, p, threads=4)
    assert np.array_equal(np.load(tmp_path / "free.npy"), x)
    assert np.array_equal(np.load(tmp_path / "trace.npy"), rep["trace_hash"])


def test_bench_rank_shards_are_disjoint_and_deterministic():
    """bench.py gives rank r the sketches [r*N, (r+1)*N) of one global seeded stream (weak scaling)."""
    from fiksi_b200 import workloads as wl
    a = wl.truss(64, first=0).raw_vars
    b = wl.truss(64, first=64).raw_vars
    both = wl.truss(128, first=0).raw_vars
    assert np.array_equal(np.vstack([a, b]), both)
    assert not np.array_equal(a, b)
