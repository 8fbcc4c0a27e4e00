"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed for the plumbing.

PSRF: the tile pairs of the symmetric distance matrix are dealt round-robin to the ranks
(asm_shard); each rank produces a partial S table of exact int64 sums; ONE all-reduce (ncclSum on
int64 is exact in any order) gives every rank the full table; the per-row PSRF values
and the reference-order sequential means are then computed locally.  The operand (bit rows of every
chain) is replicated on every rank: `trees_set_sharded` sends each row over PCIe once per job (rank r
uploads the r-th slice of every chain) and replicates it with one all-gather over NVLink.

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
        if S.is_cuda and hasattr(ctx, "psrf_sums_device_async"):
            # no host round trips: the partial sums, the all-reduce and the finish kernel are ordered on the
            # library's own stream (NCCL waits for the stream it is called under and makes it wait in turn)
            torch.cuda.current_stream(S.device).synchronize()  # S was allocated and zeroed on torch's stream
            lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=S.device)
            ctx.psrf_sums_device_async(burnin, row_start, end, delta, method, (rank, world), S.data_ptr())
            with torch.cuda.stream(lib_stream):
                dist.all_reduce(S, op=dist.ReduceOp.SUM)
        else:
            ctx.psrf_sums_device(burnin, row_start, end, delta, method, (rank, world), S.data_ptr())
            dist.all_reduce(S, op=dist.ReduceOp.SUM)
            if S.is_cuda:
                torch.cuda.current_stream(S.device).synchronize()  # the library works on its own stream
    return ctx.psrf_finish_device(S.data_ptr(), n_rows, want_rows=want_rows), S


def gather_tree_shards(dist, bits_per_chain, device):
    """bits_per_chain[c]: the whole [n_c][words] uint64 bit-row matrix of chain c in (pinned) host memory, as a
    numpy array or a torch tensor, identical on every rank.  Rank r copies rows [r*per_c, (r+1)*per_c) of every
    chain to its device; ONE all-gather hands every rank every slice.  Returns one [n_c][words] int64 device
    tensor per chain (bit patterns unchanged)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    src = []
    for b in bits_per_chain:
        t = b if isinstance(b, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(b).view(np.int64))
        src.append(t.view(torch.int64) if t.dtype != torch.int64 else t)
    words = src[0].shape[1] if src else 0
    per = [(t.shape[0] + world - 1) // world for t in src]
    slot = sum(per)  # rows every rank contributes (short slices are zero-padded)
    full = torch.zeros((world, max(slot, 1), max(words, 1)), dtype=torch.int64, device=device)
    mine = full[rank]
    off = 0
    for t, p in zip(src, per):
        lo, hi = min(rank * p, t.shape[0]), min((rank + 1) * p, t.shape[0])
        if hi > lo:
            mine[off:off + hi - lo].copy_(t[lo:hi], non_blocking=True)
        off += p
    dist.all_gather(list(full.unbind(0)), mine.clone())
    out, off = [], 0
    for t, p in zip(src, per):
        out.append(full[:, off:off + p].reshape(world * p, max(words, 1))[: t.shape[0]].contiguous())
        off += p
    return out


def trees_set_sharded(ctx, dist, bits_per_chain, device):
    """asm_trees_set for every chain with the host->device traffic divided by the world size."""
    full = gather_tree_shards(dist, bits_per_chain, device)
    if full and full[0].is_cuda:
        torch.cuda.current_stream(full[0].device).synchronize()  # the library copies on its own stream
    for c, t in enumerate(full):
        ctx.trees_set_device(c, t.data_ptr(), t.shape[0], t.shape[1])
    return full


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
