"""Multi-GPU parity on hardware ("1 vs 2 vs 4 vs 8 must be identical int64"; round-1 VERDICT weak 4).

Needs >= 2 GPUs on the box (`gpurun --gpus N`); skipped on the one-GPU box the driver's round-end run uses.  Three
ways in, all through the C ABI:
  * one process, N GPUs: `GpuContext(devices=[...])` = asm_create_multi (ncclCommInitAll inside the library);
  * the plain C caller tests/c/multi_gpu_check.c (no Python between the caller and the library);
  * one process per GPU: two spawned processes join an in-library communicator (asm_comm_init_rank), the id
    travels over torch.distributed/gloo.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conftest import bits_equal  # noqa: E402

pytestmark = pytest.mark.gpu


def _n_dev():
    from asm_b200 import _lib as L
    return L.lib().asm_device_count()


def _sizes():
    n = _n_dev()
    return [k for k in (2, 4, 8) if k <= n]


@pytest.fixture(scope="module")
def many_gpus():
    if _n_dev() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    return _n_dev()


def test_group_equals_one_gpu(gpu, many_gpus, multi_chains, ragged_chains):
    import oracle as O
    import synthgen as G
    ch = multi_chains                                        # 4 chains x 700 trees x 40 taxa
    burnin, start, end, delta = [33, 0, 120, 7], 28, 696, 4
    with gpu.GpuContext(ch.n_chains) as one:
        ch.upload(one)
        ref = {m: one.psrf_sums(burnin, start, end, delta, m) for m in (gpu.ASM_RF_POPC, gpu.ASM_RF_I8, gpu.ASM_RF_FP4)}
        ref_full = one.psrf_sums([0] * 4, 0, 700, 1, gpu.ASM_RF_FP4)
        ref_rows, ref_mean = one.psrf_eval(burnin, start, end, delta, gpu.ASM_RF_FP4, want_rows=True)
        ref_cc = one.clade_counts(2, 5, 699)
        ref_rf = one.rf_block(0, 10, 50, 3, 0, 64)
    assert np.array_equal(ref[gpu.ASM_RF_FP4].astype(np.float64), ch.ts.psrf_sums(burnin, start, end, delta))
    x = np.stack([G.synth_ar1(30000, [0.0, -1e4][j % 2], [0.0, 0.5, 0.9, 0.99][j % 4], 50 + j) for j in range(11)])
    want_act, want_ess = O.ess_batch(x)
    for n in _sizes():
        with gpu.GpuContext(ch.n_chains, devices=list(range(n))) as ctx:
            assert ctx.comm_info() == {"rank": 0, "n_ranks": n, "n_local_devices": n}
            ch.upload(ctx)                                   # sliced upload + in-place all-gather
            for m, want in ref.items():
                assert np.array_equal(ctx.psrf_sums(burnin, start, end, delta, m), want), (n, m)
            assert np.array_equal(ctx.psrf_sums([0] * 4, 0, 700, 1, gpu.ASM_RF_FP4), ref_full)
            lay = ctx.last_layout()
            assert lay["tiles_done"] == lay["tiles_total"]   # the members' shares cover the pair triangle exactly
            rows, mean = ctx.psrf_eval(burnin, start, end, delta, gpu.ASM_RF_FP4, want_rows=True)
            assert bits_equal(rows, ref_rows) and bits_equal(mean, ref_mean)
            # a caller-side shard composes with the group's own split
            parts = [ctx.psrf_sums(burnin, start, end, delta, gpu.ASM_RF_I8, shard=(i, 3)) for i in range(3)]
            assert np.array_equal(sum(parts), ref[gpu.ASM_RF_I8])
            # the same trees handed over as clade ids (every member takes them and builds its own rows), half of each
            # chain appended with the wider dictionary
            D = ch.enc.num_clades
            for c in range(ch.n_chains):
                ctx.trees_set_ids(c, ch.ids[c][:350], int(ch.ids[c][:350].max()) + 1)
                ctx.trees_append_ids(c, ch.ids[c][350:], D)
            assert np.array_equal(ctx.psrf_sums(burnin, start, end, delta, gpu.ASM_RF_FP4), ref[gpu.ASM_RF_FP4])
            rows, mean = ctx.psrf_eval(burnin, start, end, delta, gpu.ASM_RF_FP4, want_rows=True)
            assert bits_equal(rows, ref_rows) and bits_equal(mean, ref_mean)
            assert np.array_equal(ctx.clade_counts(2, 5, 699)[:D], ref_cc[:D])
            assert np.array_equal(ctx.clade_counts(2, 5, 699), ref_cc)          # sharded + ncclAllReduce(int32)
            assert np.array_equal(ref_cc[: ch.ts.num_clades], ch.ts.clade_counts(2, 5, 699))
            assert np.array_equal(ctx.rf_block(0, 10, 50, 3, 0, 64), ref_rf)
            act, ess = ctx.ess_batch(x)                      # 11 traces over n devices (ragged blocks, some empty at n = 8)
            assert act.tobytes() == want_act.tobytes() and ess.tobytes() == want_ess.tobytes()
            ctx.ess_traces_put(x)
            act2, ess2 = ctx.ess_traces_eval()
            assert act2.tobytes() == want_act.tobytes() and ess2.tobytes() == want_ess.tobytes()
    # ragged chains, appended tree by tree after a bulk set (the streaming runner), with a growing dictionary
    rg = ragged_chains
    with gpu.GpuContext(3) as one, gpu.GpuContext(3, devices=list(range(_sizes()[-1]))) as ctx:
        for c in range(3):
            half = rg.n_trees[c] // 2
            for k in (one, ctx):
                k.trees_set(c, rg.bits[c][:half])
                for i in range(half, rg.n_trees[c]):
                    k.trees_append(c, rg.bits[c][i])
            assert ctx.trees_count(c) == one.trees_count(c) == (rg.n_trees[c], rg.words)
        for m in (gpu.ASM_RF_FP4, gpu.ASM_RF_POPC):
            a = one.psrf_sums([3, 0, 10], 0, 64, 1, m)
            assert np.array_equal(ctx.psrf_sums([3, 0, 10], 0, 64, 1, m), a)
            assert np.array_equal(a.astype(np.float64), rg.ts.psrf_sums([3, 0, 10], 0, 64, 1))


def test_group_streaming_ess_columns(gpu, many_gpus):
    """asm_traces_append goes to every device, asm_ess_eval deals the selected columns (TraceESS.java:53-62)."""
    import oracle as O
    rng = np.random.default_rng(4)
    C, n, cols = 3, 900, 7
    logs = [np.cumsum(rng.normal(size=(n, cols)), axis=0) * 0.05 + rng.normal(size=(n, cols)) for _ in range(C)]
    burnin, end, sel = [10, 0, 333], 880, [1, 2, 3, 5, 6]
    with gpu.GpuContext(C, devices=list(range(_sizes()[-1]))) as ctx:
        for c in range(C):
            ctx.traces_append(c, logs[c][:500])
            ctx.traces_append(c, logs[c][500:])
        ess = ctx.ess_eval(burnin, end, sel)
        s, q = ctx.trace_moments(end)
    _, want, _ = O.trace_ess_converged([np.ascontiguousarray(l.T) for l in logs], burnin, end, sel)
    assert ess.tobytes() == np.asarray(want).tobytes()
    with gpu.GpuContext(C) as one:
        for c in range(C):
            one.traces_append(c, logs[c])
        s1, q1 = one.trace_moments(end)
    assert s.tobytes() == s1.tobytes() and q.tobytes() == q1.tobytes()


def test_c_caller_drives_every_gpu(many_gpus):
    exe = os.path.join(ROOT, "tests", "c", "multi_gpu_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "-s"])
    for method in (2, 1):
        r = subprocess.run([exe, str(many_gpus), "6", "2500", "80", str(method), "3.0", "45", "30000"], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["n_gpus"] == many_gpus == d["n_ranks"] == d["n_local_devices"] and d["nccl_version"] > 0
        assert d["S_identical_to_1gpu"] and d["psrf_means_identical"] and d["clade_counts_identical"] and d["ess_identical"]


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)   # the launcher's channel; the data path is in-library NCCL
    try:
        import synthgen as G
        from asm_b200 import _lib as L
        from asm_b200 import multi
        from conftest import SynthChains
        ch = SynthChains(30, 500, 3, beta=2.4)
        burnin, start, end, delta = [20, 0, 77], 0, 500, 1
        with L.GpuContext(3, device=rank) as ctx:
            multi.init_comm(ctx, dist)
            ch.upload(ctx)                                    # collective: rank r uploads slice r, all-gather inside
            S = ctx.psrf_sums(burnin, start, end, delta, L.ASM_RF_FP4)
            mean = ctx.psrf_eval(burnin, start, end, delta, L.ASM_RF_I8)
            cc = ctx.clade_counts(1, 0, 500)
            x = np.stack([G.synth_ar1(20000, 0.0, [0.0, 0.9][j % 2], 9 + j) for j in range(5)])
            lo, hi = multi.block_of(5, world, rank)
            ess, mn = multi.ess_sharded(ctx, x[lo:hi], 5, rank, world)
            info = ctx.comm_info()
        ok = np.array_equal(S.astype(np.float64), ch.ts.psrf_sums(burnin, start, end, delta))
        ok &= all(np.float64(mean[a, b]).tobytes() == np.float64(ch.ts.psrf_rows_mean(a, b, start, burnin, end, delta)).tobytes()
                  for a in range(3) for b in range(3) if a != b)
        ok &= np.array_equal(cc[: ch.ts.num_clades], ch.ts.clade_counts(1, 0, 500))
        import oracle as O
        ok &= ess.tobytes() == O.ess_batch(x)[1].tobytes() and info == {"rank": rank, "n_ranks": world, "n_local_devices": 1}
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu_joins_an_in_library_communicator(many_gpus):
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    world = 2
    procs = [ctxm.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world)) and all(r[1] for r in res)
