"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot
finish these sizes in seconds): symmetry of the pair sums, agreement of the two distance kernels, shard
additivity, exact invariance of ACT under power-of-two scaling, staged == all-lags."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    import bench
    enc, bits, _ = bench.gen_psrf_input(bench.CFG2, bench.CFG2["n_trees"])
    return bench.CFG2, enc, bits


def test_cfg2_full_size_properties(gpu, cfg2):
    cfg, enc, bits = cfg2
    C, N = cfg["n_chains"], cfg["n_trees"]
    with gpu.GpuContext(C) as ctx:
        for c in range(C):
            ctx.trees_set(c, bits[c])
        b0 = [0] * C
        S_i8 = ctx.psrf_sums(b0, 0, N, 1, gpu.ASM_RF_I8)
        S_pc = ctx.psrf_sums(b0, 0, N, 1, gpu.ASM_RF_POPC)
        assert np.array_equal(S_i8, S_pc)                       # tensor-core and popcount kernels agree exactly
        S_f4 = ctx.psrf_sums(b0, 0, N, 1, gpu.ASM_RF_FP4)
        assert np.array_equal(S_f4, S_i8)                       # so does the block-scaled fp4 kernel (240-row blocks)
        parts4 = [ctx.psrf_sums(b0, 0, N, 1, gpu.ASM_RF_FP4, shard=(i, 8)) for i in range(8)]
        assert np.array_equal(sum(parts4), S_i8)
        assert np.array_equal(ctx.psrf_sums([1000, 0, 37, 9999], 120, N - 3, 3, gpu.ASM_RF_FP4),
                              ctx.psrf_sums([1000, 0, 37, 9999], 120, N - 3, 3, gpu.ASM_RF_I8))  # burn-in, thinning, ragged
        tot = S_i8.sum(axis=1)                                  # [a][c] = sum over all (row in a, column in c) pairs
        assert np.array_equal(tot, tot.T)                       # RF(a,b) = RF(b,a)
        assert (S_i8[np.arange(C), :, np.arange(C)] >= 1).all()  # the diagonal term (0+1)^2
        parts = [ctx.psrf_sums(b0, 0, N, 1, gpu.ASM_RF_I8, shard=(i, 8)) for i in range(8)]
        assert np.array_equal(sum(parts), S_i8)                 # the 8-GPU split adds up exactly
        # host recomputation of three rows over all 40,000 columns (numpy XOR + popcount)
        rng = np.random.default_rng(0)
        for _ in range(3):
            a, r = int(rng.integers(C)), int(rng.integers(N))
            for c in range(C):
                d = np.bitwise_count(np.bitwise_xor(bits[c], bits[a][r][None, :])).sum(axis=1, dtype=np.int64)
                assert int(((d + 1) ** 2).sum()) == int(S_i8[a, r, c])
        # burn-in / thinning consistency: columns trimmed by burn-in only ever remove terms
        S_b = ctx.psrf_sums([1000] * C, 0, N, 1, gpu.ASM_RF_I8)
        assert (S_b <= S_i8).all() and (S_b[:, 1000:, :] < S_i8[:, 1000:, :]).all()
        rows, mean = ctx.psrf_eval(b0, 0, N, 1, gpu.ASM_RF_I8, want_rows=True)
        want = np.sqrt(S_i8[0, :, 1].astype(np.float64) / S_i8[0, :, 0].astype(np.float64))
        assert np.array_equal(rows[0, :, 1], want)
        s = 0.0
        for v in want:
            s += v
        assert mean[0, 1] == s / N                              # the reference's sequential mean


def test_cfg3_sized_ess_properties(gpu):
    import torch
    from asm_b200 import _lib as L
    import synthgen as G
    n_tr, n = 64, 1_000_000
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, [0.0, 1.0, -1e4][j % 3], [0.0, 0.5, 0.9, 0.99][j % 4], 777 + j, out=x[j])
    with gpu.GpuContext(1) as ctx:
        act, ess = ctx.ess_batch(x)
        act4, ess4 = ctx.ess_batch(x * 4.0)                     # exact in binary floating point
        assert np.array_equal(act, act4) and np.array_equal(ess, ess4)
        ctx.ess_set_adaptive(False)
        act_f, ess_f = ctx.ess_batch(x)
        assert np.array_equal(act, act_f) and np.array_equal(ess, ess_f)
        dev = torch.from_numpy(x).cuda()
        act_d, ess_d = ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n)
        assert np.array_equal(act, act_d)
    for j, phi in enumerate([0.0, 0.5, 0.9, 0.99] * 2):
        assert abs(act[j] / ((1 + phi) / (1 - phi)) - 1) < 0.2  # ACT of an AR(1) process
