"""World-size-2 test of the multi-GPU host logic on CPU (gloo).  The device side is replaced by a
stand-in that produces each rank's partial S table from the oracle, split the way asm_shard splits
tile pairs (any disjoint cover works: the sums are exact integers); what is under test is the
orchestration in asm_b200/multi.py -- shard arguments, the int64 all-reduce, the finish step, the
ESS gather and the NaN-propagating min."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StandInCtx:
    """Mimics GpuContext's device-pointer entry points on host memory."""

    def __init__(self, ts, n_chains):
        self.ts, self.n_chains = ts, n_chains
        self.calls = []

    def n_rows(self, row_start, end, delta):
        return (end - row_start + delta - 1) // delta if end > row_start else 0

    def _view(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        buf = (ctypes.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def psrf_sums_device(self, burnin, row_start, end, delta, method, shard, d_S_ptr):
        idx, cnt = shard
        self.calls.append(shard)
        C, n = self.n_chains, self.n_rows(row_start, end, delta)
        full = self.ts.psrf_sums(burnin, row_start, end, delta).astype(np.int64)
        part = np.zeros_like(full)
        # deal (row, chain) cells round-robin, like tile pairs t % shard_count == shard_index
        cells = np.arange(full.size).reshape(full.shape)
        mask = (cells % cnt) == idx
        part[mask] = full[mask]
        self._view(d_S_ptr, (C, n, C), np.int64)[...] = part

    def trees_set_device(self, chain, d_ptr, n_trees, words):
        self.trees = getattr(self, "trees", {})
        self.trees[chain] = self._view(d_ptr, (n_trees, words), np.uint64).copy()

    def psrf_finish_device(self, d_S_ptr, n_rows, want_rows=False):
        C = self.n_chains
        S = self._view(d_S_ptr, (C, max(n_rows, 1), C), np.int64)[:, :n_rows, :].astype(np.float64)
        mean = np.empty((C, C))
        for a in range(C):
            for b in range(C):
                s = 0.0
                for r in range(n_rows):  # TreePSRF.java:264-271, :287-292
                    s += np.sqrt(S[a, r, b] / S[a, r, a]) if S[a, r, a] != 0 else np.sqrt(S[a, r, b])
                mean[a, b] = s / n_rows if n_rows else float("nan")
        return mean


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from asm_b200 import multi
        from conftest import SynthChains
        ch = SynthChains(14, 60, 3, beta=1.5)
        ctx = StandInCtx(ch.ts, 3)
        burnin, start, end, delta = [4, 0, 9], 4, 60, 2
        mean, S = multi.psrf_eval_multi(ctx, dist, burnin, start, end, delta, 0, "cpu")
        want_S = ch.ts.psrf_sums(burnin, start, end, delta)
        ok = np.array_equal(S.numpy().astype(np.float64), want_S) and ctx.calls == [(rank, world)]
        for a in range(3):
            for b in range(3):
                if a != b:
                    ok &= mean[a, b] == ch.ts.psrf_rows_mean(a, b, start, burnin, end, delta)
        # sharded upload: ragged row counts (odd, smaller than the world, empty), bit patterns with the top bit set
        rng = np.random.default_rng(5)
        bits = [rng.integers(0, 2**64, size=(n, 3), dtype=np.uint64) for n in (7, 1, 0)]
        multi.trees_set_sharded(ctx, dist, bits, "cpu")
        ok &= all(np.array_equal(ctx.trees[c], bits[c]) for c in range(3))
        pinned = [torch.from_numpy(b.copy()) for b in bits]  # uint64 tensors, as bench.py passes them
        full = multi.gather_tree_shards(dist, pinned, "cpu")
        ok &= all(np.array_equal(full[c].numpy().view(np.uint64), bits[c]) for c in range(3))
        # ESS: 5 traces over 2 ranks, one of them constant (NaN)
        rng = np.random.default_rng(0)
        x = rng.normal(size=(5, 400))
        x[3] = 1.0
        mine = multi.my_traces(5, rank, world)
        ess, mn = multi.ess_multi(None, dist, x[mine], 5, "cpu", ess_fn=lambda t: O.ess_batch(np.ascontiguousarray(t)))
        want = O.ess_batch(x)[1]
        ok &= np.array_equal(ess, want, equal_nan=True) and np.isnan(mn)
        ess2, mn2 = multi.ess_multi(None, dist, x[[i for i in mine if i != 3]][: len(multi.my_traces(4, rank, world))], 4,
                                    "cpu", ess_fn=lambda t: O.ess_batch(np.ascontiguousarray(t)))
        q.put((rank, bool(ok), float(mn2)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    from asm_b200.multi import java_min
    assert np.isnan(java_min([3.0, float("nan"), 1.0])) and java_min([3.0, 1.0]) == 1.0 and java_min([]) == float("inf")
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
