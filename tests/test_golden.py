"""Committed fixtures (tests/golden/, produced by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must reproduce them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

import oracle as O
from conftest import bits_equal

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def psrf():
    return np.load(os.path.join(G, "psrf_small.npz"))


@pytest.fixture(scope="module")
def ess():
    return np.load(os.path.join(G, "ess_small.npz"))


def _treeset(g):
    C, n_taxa = int(g["n_chains"]), int(g["n_taxa"])
    ts = O.TreeSet(C, n_taxa)
    for c in range(C):
        ts.add(c, g[f"left{c}"], g[f"right{c}"], g[f"root{c}"])
    return ts


def test_oracle_reproduces_psrf_golden(psrf):
    ts = _treeset(psrf)
    assert ts.num_clades == int(psrf["num_clades"])
    assert np.array_equal(ts.rf_block(0, 0, 90, 2, 0, 90), psrf["rf_0_2"])
    for c in range(3):
        assert np.array_equal(ts.clade_counts(c, 5, 80), psrf[f"clade_counts{c}"])
    for k, case in enumerate(psrf["cases"]):
        b, (s, e, d) = case[:3].tolist(), case[3:].tolist()
        assert np.array_equal(ts.psrf_sums(b, s, e, d), psrf[f"S{k}"].astype(np.float64))
        for a in range(3):
            for bb in range(3):
                if a != bb:
                    assert bits_equal(np.array([ts.psrf_rows_mean(a, bb, s, b, e, d)]), psrf[f"mean{k}"][a, bb:bb + 1])


def test_oracle_reproduces_ess_golden(ess):
    for n in ess["lens"]:
        act, e = O.ess_batch(ess[f"x{n}"])
        assert bits_equal(act, ess[f"act{n}"]) and bits_equal(e, ess[f"ess{n}"])


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
def test_gpu_reproduces_psrf_golden(gpu, psrf, method):
    C = int(psrf["n_chains"])
    with gpu.GpuContext(C) as ctx:
        for c in range(C):
            ctx.trees_set(c, psrf[f"bits{c}"])
        assert np.array_equal(ctx.rf_block(0, 0, 90, 2, 0, 90, method), psrf["rf_0_2"])
        D = int(psrf["num_clades"])
        for c in range(C):
            assert np.array_equal(ctx.clade_counts(c, 5, 80)[:D], psrf[f"clade_counts{c}"])
        for k, case in enumerate(psrf["cases"]):
            b, (s, e, d) = case[:3].tolist(), case[3:].tolist()
            assert np.array_equal(ctx.psrf_sums(b, s, e, d, method), psrf[f"S{k}"])
            rows, mean = ctx.psrf_eval(b, s, e, d, method, want_rows=True)
            off = ~np.eye(C, dtype=bool)
            assert bits_equal(mean[off], psrf[f"mean{k}"][off])
            for a in range(C):
                for bb in range(C):
                    if a != bb:
                        assert bits_equal(rows[a, :, bb], psrf[f"rows{k}"][a, :, bb])


@pytest.mark.gpu
def test_gpu_reproduces_ess_golden(gpu, ess):
    with gpu.GpuContext(1) as ctx:
        for n in ess["lens"]:
            act, e = ctx.ess_batch(ess[f"x{n}"])
            assert bits_equal(act, ess[f"act{n}"]) and bits_equal(e, ess[f"ess{n}"])
