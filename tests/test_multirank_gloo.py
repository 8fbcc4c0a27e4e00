"""World-size-2 test of the launcher-side multi-GPU logic on CPU (gloo): what asm_b200/multi.py adds around the
library for the one-process-per-GPU mode -- carrying the communicator id from rank 0 to the other ranks, the
contiguous block dealing of ESS traces, reassembling the per-rank ESS vectors, the NaN-propagating min.  The device
side (NCCL inside libasmgpu.so) is replaced by a stand-in whose all-gather runs over gloo; the collectives
themselves are tested on hardware by tests/test_gpu_multi.py and tests/c/multi_gpu_check.c."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StandInCtx:
    """GpuContext's communicator-facing calls on host memory (ESS through the oracle, all-gather over gloo)."""

    def __init__(self):
        self.joined = None

    def comm_init_rank(self, comm_id, rank, n_ranks):
        self.joined = (bytes(comm_id), rank, n_ranks)

    def ess_batch(self, traces):
        import oracle as O
        t = np.ascontiguousarray(traces, dtype=np.float64)
        if t.shape[0] == 0:
            return np.empty(0), np.empty(0)
        return O.ess_batch(t)

    def all_gather_f64(self, local):
        world = dist.get_world_size()
        out = [torch.empty(len(local), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)))
        return torch.cat(out).numpy()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from asm_b200 import multi
        ctx = StandInCtx()
        cid = multi.init_comm(ctx, dist, make_id=lambda: bytes(range(128)))
        ok = ctx.joined == (bytes(range(128)), rank, world) and cid == bytes(range(128))
        # ESS: 5 traces over 2 ranks (blocks of 3 and 2), one of them constant (NaN)
        rng = np.random.default_rng(0)
        x = rng.normal(size=(5, 400))
        x[3] = 1.0
        lo, hi = multi.block_of(5, world, rank)
        ess, mn = multi.ess_sharded(ctx, x[lo:hi], 5, rank, world)
        want = O.ess_batch(x)[1]
        ok &= np.array_equal(ess, want, equal_nan=True) and np.isnan(mn)
        # fewer traces than ranks would leave a rank empty: 1 trace over 2 ranks
        lo1, hi1 = multi.block_of(1, world, rank)
        ess1, mn1 = multi.ess_sharded(ctx, x[lo1:hi1], 1, rank, world)
        ok &= np.array_equal(ess1, want[:1]) and mn1 == want[0]
        q.put((rank, bool(ok), float(mn1)))
    finally:
        dist.destroy_process_group()


def test_block_dealing_covers_everything_once():
    from asm_b200.multi import block_of, java_min
    for n in (0, 1, 5, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 4, 8):
            blocks = [block_of(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(hi - lo for lo, hi in blocks) == (n + world - 1) // world
    assert np.isnan(java_min([3.0, float("nan"), 1.0])) and java_min([3.0, 1.0]) == 1.0 and java_min([]) == float("inf")


def test_two_rank_gloo():
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
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res) and res[0][2] == res[1][2]
