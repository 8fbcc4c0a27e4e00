"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed for the plumbing.

PSRF: the tile pairs of the symmetric distance matrix are dealt round-robin to the ranks
(asm_shard); each rank produces a partial S table of exact int64 sums; ONE all-reduce (ncclSum on
int64 is exact in any order) gives every rank the full table; the per-row PSRF values
and the reference-order sequential means are then computed locally.  The operand (bit rows of every
chain) is replicated on every rank.

ESS: traces are independent, so trace j goes to rank j % world; the ESS vector is all-gathered and
the min is taken on the host with Java's NaN-propagating Math.min (TraceESS.java:63) -- NCCL's
min would drop NaNs.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch


def java_min(values) -> float:
    """Math.min folded over `values` starting from +inf: NaN if any element is NaN (TraceESS.java:50,63)."""
    m = math.inf
    for v in values:
        v = float(v)
        if math.isnan(v) or math.isnan(m):
            m = math.nan
        else:
            m = min(m, v)
    return m


def psrf_eval_multi(ctx, dist, burnin: Sequence[int], row_start: int, end: int, delta: int, method: int,
                    device, want_rows: bool = False):
    """Every rank returns the same psrf_mean [C][C] (and rows if asked)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    C = ctx.n_chains
    n_rows = ctx.n_rows(row_start, end, delta)
    S = torch.zeros((C, max(n_rows, 1), C), dtype=torch.int64, device=device)
    if n_rows > 0:
        ctx.psrf_sums_device(burnin, row_start, end, delta, method, (rank, world), S.data_ptr())
        dist.all_reduce(S, op=dist.ReduceOp.SUM)
        if S.is_cuda:
            torch.cuda.current_stream(S.device).synchronize()  # the library works on its own stream
    return ctx.psrf_finish_device(S.data_ptr(), n_rows, want_rows=want_rows), S


def my_traces(n_traces: int, rank: int, world: int):
    return list(range(rank, n_traces, world))


def ess_multi(ctx, dist, traces_local, n_traces_total: int, device, ess_fn=None):
    """traces_local: the rows my_traces(...) of the full [n_traces][n] matrix, in that order.
    Returns (ess[n_traces_total], minESS) on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = my_traces(n_traces_total, rank, world)
    _, ess_local = (ess_fn or ctx.ess_batch)(traces_local)
    per_rank = (n_traces_total + world - 1) // world
    buf = torch.full((per_rank,), float("nan"), dtype=torch.float64, device=device)
    buf[: len(mine)] = torch.from_numpy(np.asarray(ess_local, dtype=np.float64)).to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    ess = np.empty(n_traces_total, dtype=np.float64)
    for r in range(world):
        idx = my_traces(n_traces_total, r, world)
        ess[idx] = gathered[r][: len(idx)].cpu().numpy()
    return ess, java_min(ess)
