"""Launcher-side helpers for ONE PROCESS PER GPU (torchrun).  The collectives themselves are inside libasmgpu.so
(asm_b200/csrc/comm.cu: NCCL all-reduce of the int64 PSRF sums and int32 clade counts, in-place all-gather of the
uploaded bit rows); what a launcher has to add is small:

  * `init_comm`: rank 0 draws the communicator id (asm_comm_unique_id), the launcher's own channel
    (torch.distributed here) carries its 128 bytes to the other ranks, every rank joins with its plain context
    (asm_comm_init_rank).  From then on `ctx.trees_set`, `ctx.psrf_sums`, `ctx.psrf_eval`, `ctx.clade_counts`
    are collective calls that return the complete result on every rank.
  * ESS: traces are independent, so every rank evaluates the block of traces it holds (`block_of`, the same
    contiguous dealing asm_create_multi uses inside one process) and `ess_sharded` combines the per-rank vectors
    with the library's all-gather; the min is taken on the host with Java's NaN-propagating Math.min
    (TraceESS.java:63) -- an NCCL min would drop NaNs.

A single process that drives several GPUs needs none of this: `GpuContext(n_chains, devices=[0, 1, ...])`.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import numpy as np


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


def block_of(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the items rank `rank` of `world` holds: contiguous blocks of ceil(n / world) (comm.cu deals
    traces, selected columns, tree slices and clade-count ranges the same way)."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def init_comm(ctx, dist, make_id: Optional[Callable[[], bytes]] = None) -> bytes:
    """Join `ctx` to an in-library NCCL communicator spanning the ranks of `dist`.  Returns the id."""
    from . import _lib
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [(make_id or _lib.comm_unique_id)() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init_rank(box[0], rank, world)
    return box[0]


def ess_sharded(ctx, traces_local: np.ndarray, n_traces_total: int, rank: int, world: int, resident: bool = False):
    """traces_local: rows block_of(n_traces_total, world, rank) of the full [n_traces][n] matrix (ignored when
    `resident`: the traces were given to asm_ess_traces_put before).  Returns (ess[n_traces_total], minESS), the same
    on every rank."""
    lo, hi = block_of(n_traces_total, world, rank)
    per = (n_traces_total + world - 1) // world
    _, ess_local = ctx.ess_traces_eval() if resident else ctx.ess_batch(traces_local)
    assert len(ess_local) == hi - lo
    buf = np.full(per, np.nan)
    buf[: hi - lo] = ess_local
    gathered = np.asarray(ctx.all_gather_f64(buf)).reshape(world, per)
    ess = np.concatenate([gathered[r][: block_of(n_traces_total, world, r)[1] - block_of(n_traces_total, world, r)[0]]
                          for r in range(world)])
    return ess, java_min(ess)
