"""Sketch-sharded batches across ranks (one process per GPU, torch.distributed for the plumbing).

Independent sketches are the only parallel axis of the path: rank g solves the
contiguous range [g*N/G, (g+1)*N/G) and nothing crosses GPUs during the solve; one gather of the
free values and reports at the end (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import REPORT_DTYPE


def shard_bounds(n: int, world: int):
    """Contiguous, balanced ranges — the same split fk_batch_solve uses across devices."""
    return [(g * n // world, (g + 1) * n // world) for g in range(world)]


def solve_sharded(solver, vars_, param, n_free, device=None, gather=True):
    """Each rank solves its shard with `solver(vars, param) -> (free[n][n_free], reports[n])`
    (normally `Topology.batch_solve`), then the shards are gathered on every rank.

    Returns (free, reports) for the whole batch if gather else for the local shard."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = vars_.shape[0]
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank]
    x, rep = solver(vars_[lo:hi], param[lo:hi])
    if not gather or world == 1:
        return x, rep
    cap = max(h - l for l, h in bounds)
    dev = device if device is not None else torch.device("cpu")
    xb = torch.zeros((cap, n_free), dtype=torch.float64, device=dev)
    rb = torch.zeros((cap, REPORT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    xb[: hi - lo] = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    rb[: hi - lo] = torch.from_numpy(np.ascontiguousarray(rep).view(np.uint8).reshape(hi - lo, -1)).to(dev)
    xs = [torch.empty_like(xb) for _ in range(world)]
    rs = [torch.empty_like(rb) for _ in range(world)]
    dist.all_gather(xs, xb)
    dist.all_gather(rs, rb)
    free = np.concatenate([xs[g][: h - l].cpu().numpy() for g, (l, h) in enumerate(bounds)])
    reports = np.concatenate([rs[g][: h - l].cpu().numpy().reshape(-1).view(REPORT_DTYPE) for g, (l, h) in enumerate(bounds)])
    return free, reports
