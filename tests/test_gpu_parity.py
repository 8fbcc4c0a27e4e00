"""GPU parity: every C-ABI compute entry point against the CPU oracle on the same seeded inputs.
Integer results (RF, clade counts, S) must be identical; PSRF/ESS doubles must be bit-identical
(which is stricter than the 1e-12 relative tolerance the north star states)."""
import numpy as np
import pytest

import oracle as O
from conftest import SynthChains, bits_equal

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12  # north-star tolerance for PSRF / ESS doubles; the tests below demand bit equality

METHODS = [0, 1, 2]  # ASM_RF_POPC, ASM_RF_I8, ASM_RF_FP4
BLOCK_METHODS = [0, 1]  # the fp4 kernel has no distance dump


def _ctx(gpu, chains):
    ctx = gpu.GpuContext(chains.n_chains)
    chains.upload(ctx)
    return ctx


# ---------------------------------------------------------------- RF distances (TreeESS.java:179-180)
@pytest.mark.parametrize("method", BLOCK_METHODS)
def test_rf_block_bit_exact(gpu, multi_chains, method):
    ch = multi_chains
    with _ctx(gpu, ch) as ctx:
        for (a, r0, r1, b, c0, c1) in [(0, 0, 300, 0, 0, 300), (0, 17, 200, 3, 5, 411), (2, 650, 700, 1, 0, 700)]:
            got = ctx.rf_block(a, r0, r1, b, c0, c1, method)
            want = ch.ts.rf_block(a, r0, r1, b, c0, c1)
            assert got.dtype == np.uint16 and np.array_equal(got, want)
        assert ctx.rf_block(0, 5, 5, 1, 0, 10, method).shape == (0, 10)  # empty range


def test_rf_known_answers(gpu):
    # two fixed 4-taxon trees ((0,1),(2,3)) and ((0,2),(1,3)): clades {01,23,0123} vs {02,13,0123} -> RF 4;
    # caterpillars (((0,1),2),3) vs (((0,2),1),3): {01,012,0123} vs {02,012,0123} -> RF 2
    import asm_b200 as A
    enc = A.Encoder(4)
    t = [enc.add_newick(s, 1) for s in ("((1,2),(3,4));", "((1,3),(2,4));", "(((1,2),3),4);", "(((1,3),2),4);",
                                        "((1:0.1,2:0.2)[&x=1]:0.3,(3:1e-2,4:1):0.5):0.0;")]
    bits = enc.bits(np.stack(t))
    with gpu.GpuContext(1) as ctx:
        ctx.trees_set(0, bits)
        d = ctx.rf_block(0, 0, 5, 0, 0, 5)
    assert d[0, 1] == 4 and d[2, 3] == 2 and d[0, 4] == 0 and np.array_equal(d, d.T) and not d.diagonal().any()


# ---------------------------------------------------------------- clade counts (CladeDifference.java:147-159)
def test_clade_counts_bit_exact(gpu, multi_chains):
    ch = multi_chains
    D = ch.ts.num_clades
    assert D == ch.enc.num_clades
    with _ctx(gpu, ch) as ctx:
        for c, (lo, hi) in enumerate([(0, 700), (100, 101), (3, 699), (0, 0)]):
            got = ctx.clade_counts(c, lo, hi)
            want = ch.ts.clade_counts(c, lo, hi)
            assert np.array_equal(got[:D], want) and not got[D:].any()
    # identity of the dictionary: same clade <-> same id on both sides
    for cid in [0, 1, D // 2, D - 1]:
        assert ch.ts.clade_key(cid) == "".join(f"{x}," for x in ch.enc.clade_leaves(cid))


# ---------------------------------------------------------------- S sums (TreePSRF.java:275-286)
PSRF_CASES = [
    # burnin, row_start, end, delta
    ([0, 0], 0, 300, 1),
    ([30, 30], 0, 300, 1),        # columns trimmed, rows not (TreePSRF.java:141-144 vs :276-277)
    ([31, 17], 30, 299, 1),       # ragged burn-in, smoothing < 1
    ([0, 0], 0, 300, 2),          # thinning
    ([33, 10], 28, 296, 4),       # burn-in rounded up to a multiple of delta (:296-301)
    ([300, 0], 0, 300, 1),        # chain 0 has no columns: varIn == 0 branch (:287-292)
    ([0, 0], 0, 1, 1),            # a single tree
]


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("burnin,row_start,end,delta", PSRF_CASES)
def test_psrf_sums_two_chains(gpu, small_chains, method, burnin, row_start, end, delta):
    ch = small_chains
    want = ch.ts.psrf_sums(burnin, row_start, end, delta)
    with _ctx(gpu, ch) as ctx:
        got = ctx.psrf_sums(burnin, row_start, end, delta, method)
    assert got.dtype == np.int64 and got.shape == want.shape
    assert np.array_equal(got.astype(np.float64), want)  # exact integers < 2^53


@pytest.mark.parametrize("method", METHODS)
def test_psrf_eval_two_chains_matches_reference_rows(gpu, small_chains, method):
    """psrf_mean[0][1] / [1][0] are psrf1mean / psrf2mean of converged2sided (TreePSRF.java:145,157)."""
    ch = small_chains
    for burnin, row_start, end, delta in PSRF_CASES:
        with _ctx(gpu, ch) as ctx:
            rows, mean = ctx.psrf_eval(burnin, row_start, end, delta, method, want_rows=True)
        for (a, b) in [(0, 1), (1, 0)]:
            if (end - row_start) % delta:  # the Java only ever calls with both divisible by delta
                continue
            m, r = ch.ts.psrf_rows_mean(a, b, row_start, burnin, end, delta, want_rows=True)
            assert bits_equal(rows[a, :, b], r), (burnin, row_start, end, delta, a, b)
            assert bits_equal(np.array([mean[a, b]]), np.array([m]))
            if np.isfinite(m) and m != 0:
                assert abs(mean[a, b] - m) <= REL_TOL * abs(m)


@pytest.mark.parametrize("method", METHODS)
def test_psrf_multi_chain(gpu, multi_chains, method):
    ch = multi_chains
    cases = [([0, 0, 0, 0], 0, 700, 1), ([70, 10, 0, 33], 70, 700, 1), ([5, 9, 300, 2], 0, 696, 8)]
    for burnin, row_start, end, delta in cases:
        want = ch.ts.psrf_sums(burnin, row_start, end, delta)
        with _ctx(gpu, ch) as ctx:
            got = ctx.psrf_sums(burnin, row_start, end, delta, method)
            rows, mean = ctx.psrf_eval(burnin, row_start, end, delta, method, want_rows=True)
        assert np.array_equal(got.astype(np.float64), want)
        for a in range(4):
            for b in range(4):
                if a == b:
                    continue
                m, r = ch.ts.psrf_rows_mean(a, b, row_start, burnin, end, delta, want_rows=True)
                assert bits_equal(rows[a, :, b], r) and bits_equal(np.array([mean[a, b]]), np.array([m]))


@pytest.mark.parametrize("method", METHODS)
def test_psrf_ragged_and_empty(gpu, ragged_chains, method):
    ch = ragged_chains
    with _ctx(gpu, ch) as ctx:
        want = ch.ts.psrf_sums([3, 0, 60], 0, 64, 1)
        assert np.array_equal(ctx.psrf_sums([3, 0, 60], 0, 64, 1, method).astype(np.float64), want)
        # end == 0: no rows, mean = 0/0 = NaN (AutoStopMCMC calls with end = 0 first)
        mean = ctx.psrf_eval([0, 0, 0], 0, 0, 1, method)
        assert np.isnan(mean).all()
        # end past the shortest chain: the Java throws inside converged2sided; here ASM_E_RANGE
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.psrf_eval([0, 0, 0], 0, 65, 1, method)
        assert e.value.code == gpu.ASM_E_RANGE
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.psrf_eval([0, 0, 0], 3, 64, 2, method)  # row_start not a multiple of delta
        assert e.value.code == gpu.ASM_E_INVALID


@pytest.mark.parametrize("method", METHODS)
def test_psrf_shards_add_up(gpu, multi_chains, method):
    """Multi-GPU split: partial S tables of the shards sum (exact int64) to the full table."""
    ch = multi_chains
    burnin, row_start, end, delta = [70, 10, 0, 33], 70, 700, 1
    with _ctx(gpu, ch) as ctx:
        full = ctx.psrf_sums(burnin, row_start, end, delta, method)
        for n in (2, 3, 8):
            parts = [ctx.psrf_sums(burnin, row_start, end, delta, method, shard=(i, n)) for i in range(n)]
            assert np.array_equal(sum(parts), full)
            assert any(p.any() for p in parts[1:])


def test_identical_trees_known_answer(gpu):
    """All trees identical: every (d+1)^2 = 1, so psrf = sqrt(m2/m1) (known answers)."""
    ch = SynthChains(12, 50, 2, beta=50.0, moves=0)
    with _ctx(gpu, ch) as ctx:
        S = ctx.psrf_sums([10, 0], 0, 50, 1)
        assert (S[:, :, 0] == 40).all() and (S[:, :, 1] == 50).all()
        mean = ctx.psrf_eval([10, 0], 0, 50, 1)
    # TreePSRF.mean (TreePSRF.java:264-271) adds the 50 identical row values one by one, then divides
    def seq_mean(v, n):
        s = 0.0
        for _ in range(n):
            s += v
        return s / n
    assert mean[0, 1] == seq_mean(np.sqrt(50.0 / 40.0), 50) and mean[1, 0] == seq_mean(np.sqrt(40.0 / 50.0), 50)
    assert abs(mean[0, 1] - np.sqrt(1.25)) < 1e-14


def test_streaming_append_and_dictionary_growth(gpu, small_chains):
    """One tree per chain per check (AutoStopMCMC.java:355-375), dictionary width growing on the way."""
    ch = small_chains
    import asm_b200 as A
    enc = A.Encoder(ch.n_taxa)
    with gpu.GpuContext(2) as ctx:
        for i in range(0, 120):
            for c in range(2):
                l, r, root = ch.topo[c]
                ids = enc.add_topologies(l[i:i + 1], r[i:i + 1], root[i:i + 1])
                ctx.trees_append(c, enc.bits(ids))  # enc.words grows in steps of 4 words
            if i in (0, 1, 7, 63, 119):
                end = i + 1
                want = ch.ts.psrf_sums([0, 0], 0, end, 1)
                got = ctx.psrf_sums([0, 0], 0, end, 1)
                assert np.array_equal(got.astype(np.float64), want)
        assert ctx.trees_count(0)[0] == 120


# ---------------------------------------------------------------- ESS (TraceESS.java:126-179)
def _traces(n_tr, n, seed=777):
    from asm_b200 import _lib as L
    import synthgen as G
    phis, mus = [0.0, 0.5, 0.9, 0.99], [0.0, 1.0, -1e4]
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, mus[j % 3], phis[j % 4], seed + j, out=x[j])
    return x


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 50, 300, 777, 1023, 1024, 1025, 1999, 2000, 2001, 2003, 4100, 10007])
def test_ess_batch_bit_exact(gpu, n):
    x = _traces(6, n)
    want_act, want_ess = O.ess_batch(x)
    with gpu.GpuContext(1) as ctx:
        act, ess = ctx.ess_batch(x)
    assert bits_equal(act, want_act) and bits_equal(ess, want_ess)
    if n <= 2003:  # the literal O(n*2000) Java loop, not just its closed form
        lit = np.array([O.act_literal(x[j]) for j in range(6)])
        assert bits_equal(act, lit)


def test_ess_staged_equals_all_lags_and_wide_path(gpu):
    """Staged (adaptive) lag evaluation must be bit-identical to evaluating every lag, on both thread shapes
    (R = 4 for few open chains, R = 8 for many)."""
    from asm_b200 import _lib as L
    import synthgen as G
    n_tr, n = 2200, 2600
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, [0.0, -1e4][j % 2], [0.0, 0.9, 0.99, 0.999][j % 4], 4000 + j, out=x[j])
    want_act, want_ess = O.ess_batch(x)
    with gpu.GpuContext(1) as ctx:
        act_a, ess_a = ctx.ess_batch(x)
        slots_a = ctx.ess_last_lag_count()
        ctx.ess_set_adaptive(False)
        act_f, ess_f = ctx.ess_batch(x)
        slots_f = ctx.ess_last_lag_count()
    assert bits_equal(act_a, want_act) and bits_equal(ess_a, want_ess)
    assert bits_equal(act_f, want_act) and bits_equal(ess_f, want_ess)
    assert slots_f == n_tr * 2048 and n_tr * 128 <= slots_a < slots_f


@pytest.mark.parametrize("lags_per_thread", [8, 4, 2, 1])
@pytest.mark.parametrize("n", [5003, 700])
def test_ess_every_thread_shape(gpu, lags_per_thread, n, monkeypatch):
    """Each lags-per-thread instantiation of the lag kernel (normally picked by the cost model) on its own,
    on a ragged length, staged and unstaged."""
    from asm_b200 import _lib as L
    import synthgen as G
    n_tr = 37
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, [0.0, 1.0, -1e4][j % 3], [0.0, 0.5, 0.9, 0.99, 0.999][j % 5], 9100 + j, out=x[j])
    want_act, want_ess = O.ess_batch(x)
    monkeypatch.setenv("ASM_ESS_R", ",".join([str(lags_per_thread)] * 5))
    with gpu.GpuContext(1) as ctx:
        for adaptive in (True, False):
            ctx.ess_set_adaptive(adaptive)
            act, ess = ctx.ess_batch(x)
            assert bits_equal(act, want_act) and bits_equal(ess, want_ess)


@pytest.mark.parametrize("plan", ["20:5,10:3,8:2,4:1", "1:5,41:2", "30:1,12:5"])
def test_ess_host_upload_groups_and_stage_plans(gpu, plan, monkeypatch):
    """The host-pointer entry uploads in chunks, each a group on its own stream with its own stage plan
    (five / three / two stages or all lags at once), served by a polling host loop.  Normally that engages
    only for gigabyte inputs; ASM_ESS_H2D_PLAN forces it on a small one."""
    from asm_b200 import _lib as L
    import synthgen as G
    n_tr, n = 42, 5003
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, [0.0, -1e4][j % 2], [0.0, 0.5, 0.9, 0.99, 0.999][j % 5], 8800 + j, out=x[j])
    want_act, want_ess = O.ess_batch(x)
    monkeypatch.setenv("ASM_ESS_H2D_PLAN", plan)
    with gpu.GpuContext(1) as ctx:
        act, ess = ctx.ess_batch(x)
        act_s, ess_s = ctx.ess_batch(np.ascontiguousarray(np.pad(x, ((0, 0), (0, 5))))[:, :n])  # row stride > n
    assert bits_equal(act, want_act) and bits_equal(ess, want_ess)
    assert bits_equal(act_s, want_act) and bits_equal(ess_s, want_ess)


def test_ess_narrowed_stage(gpu):
    """More open traces in a late stage than one wave of warps holds (148 SMs x 4 sub-partitions): the stage is
    narrowed to a single wave and the next one picks up the rest.  Same bits as the oracle and as all lags."""
    from asm_b200 import _lib as L
    import synthgen as G
    n_tr, n = 130, 2500
    x = np.empty((n_tr, n))
    for j in range(n_tr):
        G.synth_ar1(n, 0.0, 0.9995 if j < 100 else 0.3, 7300 + j, out=x[j])
    want_act, want_ess = O.ess_batch(x)
    with gpu.GpuContext(1) as ctx:
        act, ess = ctx.ess_batch(x)
        used = ctx.ess_last_lag_count()
        ctx.ess_set_adaptive(False)
        act_f, ess_f = ctx.ess_batch(x)
    assert bits_equal(act, want_act) and bits_equal(ess, want_ess)
    assert bits_equal(act_f, want_act) and bits_equal(ess_f, want_ess)
    assert used < n_tr * 2048


def test_ess_edge_cases(gpu):
    with gpu.GpuContext(1) as ctx:
        act, ess = ctx.ess_batch(np.full((2, 100), 3.25))  # constant trace: ac[0] = 0 -> NaN
        assert np.isnan(act).all() and np.isnan(ess).all()
        act, ess = ctx.ess_batch(np.zeros((3, 0)))  # n = 0: 0/0
        assert np.isnan(act).all() and np.isnan(ess).all()
        x = _traces(4, 50000)  # iid / AR(1): ACT ~ (1+phi)/(1-phi)
        act, _ = ctx.ess_batch(x)
        for j, phi in enumerate([0.0, 0.5, 0.9, 0.99]):
            assert abs(act[j] / ((1 + phi) / (1 - phi)) - 1) < 0.35
        # strided input (row stride > n)
        big = _traces(3, 1000)
        act2, _ = ctx.ess_batch(big[:, :600])
        assert bits_equal(act2, O.ess_batch(np.ascontiguousarray(big[:, :600]))[0])
        with pytest.raises(gpu.AsmGpuError):
            ctx.ess_batch(x, max_lag=0)


def test_ess_eval_streaming_concat(gpu):
    """Chain concatenation of TraceESS.converged (TraceESS.java:53-62) done on the device."""
    n_chains, n_cols, n = 3, 5, 900
    rng = np.random.default_rng(5)
    logs = [np.cumsum(rng.normal(size=(n_cols, n)), axis=1) * 0.1 + rng.normal(size=(n_cols, n)) for _ in range(n_chains)]
    with gpu.GpuContext(n_chains) as ctx:
        for i in range(0, n, 100):  # appended in blocks of log lines [n_new][n_cols]
            for c in range(n_chains):
                ctx.traces_append(c, np.ascontiguousarray(logs[c][:, i:i + 100].T))
        for burnin, end, cols in [([0, 0, 0], 900, [1, 2, 3]), ([90, 10, 450], 700, [4, 0]), ([5, 5, 5], 5, [1])]:
            got = ctx.ess_eval(burnin, end, cols)
            conv, want, mn = O.trace_ess_converged(logs, burnin, end, cols, targetESS=100)
            assert bits_equal(got, want)
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.ess_eval([0, 0, 0], 901, [1])
        assert e.value.code == gpu.ASM_E_RANGE


def test_errors_are_codes_not_crashes(gpu, small_chains):
    with gpu.GpuContext(2) as ctx:
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.trees_set(5, small_chains.bits[0])
        assert e.value.code == gpu.ASM_E_INVALID
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.rf_block(0, 0, 10, 1, 0, 10)  # nothing stored yet
        assert e.value.code == gpu.ASM_E_RANGE
    with pytest.raises(gpu.AsmGpuError):
        gpu.GpuContext(0)


# ---------------------------------------------------------------- GelmanRubin / CladeDifference
def test_gelman_rubin_and_clade_difference_criteria(gpu, multi_chains):
    from asm_b200 import criteria as K
    ch = multi_chains
    rng = np.random.default_rng(11)
    n, n_cols = 700, 5
    logs = [np.cumsum(rng.normal(size=(n_cols, n)), axis=1) * 0.02 + rng.normal(size=(n_cols, n)) + 0.3 * c for c in range(4)]
    ti = K.TraceInfo(4)
    ti.setTaxonCount(ch.n_taxa)
    for c in range(4):
        l, r, root = ch.topo[c]
        for i in range(n):
            ti.addTree(c, l[i], r[i], int(root[i]))
            ti.addLogLine(c, logs[c][:, i])
    gr, cd = K.GelmanRubin(threshold=1.05), K.CladeDifference(threshold=0.25)
    gr.setup(4, ti)
    cd.setup(4, ti)
    for available in (0, 1, 2, 50, 333, 700):
        want, stats = O.gelman_rubin_converged(logs, available, 1.05)
        assert gr.converged([0, 0, 0, 0], available) == want, available
        if available > 2:
            assert bits_equal(np.array([gr.lastValue]), np.array([max(0.0, np.nanmax(stats))]))
        want_c, mx = O.clade_difference_converged(ch.ts, available, 0.25)
        assert cd.converged([0, 0, 0, 0], available) == want_c, available
        assert bits_equal(np.array([cd.lastValue]), np.array([mx]))
    with pytest.raises(K.IllegalArgumentException):
        gr.converged([0, 0, 0, 0], 701)   # "Expected traces of sufficient length" (GelmanRubin.java:56-58)
    # raw moments, Java order
    with gpu.GpuContext(2) as ctx:
        for c in range(2):
            ctx.traces_append(c, np.ascontiguousarray(logs[c].T))
        s, q = ctx.trace_moments(500)
        for c in range(2):
            for k in range(n_cols):
                ws = wq = 0.0
                for d in logs[c][k, :500]:
                    ws += d
                    wq += d * d
                assert s[c, k] == ws and q[c, k] == wq


@pytest.mark.gpu
def test_deliberate_deviations_of_the_host_mirror(gpu, small_chains):
    """The places where the C++ mirror does NOT do what the Java does ("Deviations"), pinned so that they
    stay deliberate."""
    from asm_b200 import criteria as K
    ch = small_chains
    rng = np.random.default_rng(5)
    # (1) TraceInfo.getMap with labels set and setUpMap never called: the reference dereferences tracesString == null
    # (TraceInfo.java:57 -> :76, NullPointerException); the mirror reads the default "posterior,likelihood,prior".
    ti = K.TraceInfo(2)
    assert ti.getMap() == [1, 2, 3]                      # no labels yet: DEFAULT_MAP (:46), as in the reference
    ti.readLogLine(0, "Sample\tprior\tposterior\tkappa\tlikelihood")
    assert ti.getMap() == [2, 4, 1]
    # (2) GelmanRubin with a single chain: no pair, true whatever `available` is -- same as the reference (:40-41)
    one = K.TraceInfo(1)
    one.addLogLine(0, [0.0, 1.0, 2.0])
    g1 = K.GelmanRubin()
    g1.setup(1, one)
    assert g1.converged([0], 1) is True and g1.converged([0], 10**6) is True
    # (3) GelmanRubin validates `available` against every chain before the first pair; the reference only throws when
    # calcGRStat reaches a short chain (:56-58), so with chains 0/1 already above the threshold it returns false
    # instead.  The runner never asks for more samples than the shortest chain holds (AutoStopMCMC.java:389-395).
    t3 = K.TraceInfo(3)
    for i in range(50):
        t3.addLogLine(0, [float(i), rng.normal()])
        t3.addLogLine(1, [float(i), 100.0 + rng.normal()])
        if i < 10:
            t3.addLogLine(2, [float(i), rng.normal()])
    g3 = K.GelmanRubin(threshold=1.05)
    g3.setup(3, t3)
    assert g3.converged([0, 0, 0], 10) is False          # enough samples everywhere: chains 0 and 1 disagree
    with pytest.raises(K.IllegalArgumentException):
        g3.converged([0, 0, 0], 50)                      # the reference: false (pair (0,1) decides first)
    # (4) CladeDifference counts [0, available) afresh on every call; the reference keeps counting from `current`
    # (CladeDifference.java:47-53), so the two differ only if `available` ever decreases, which the runner never does.
    tc = K.TraceInfo(2)
    tc.setTaxonCount(ch.n_taxa)
    for c in range(2):
        l, r, root = ch.topo[c]
        for i in range(300):
            tc.addTree(c, l[i], r[i], int(root[i]))
    cd = K.CladeDifference(threshold=0.25)
    cd.setup(2, tc)
    cd.converged([0, 0], 300)
    cd.converged([0, 0], 120)                            # the reference would still report the 300-tree counts
    _, mx120 = O.clade_difference_converged(ch.ts, 120, 0.25)
    assert bits_equal(np.array([cd.lastValue]), np.array([mx120]))


@pytest.mark.gpu
@pytest.mark.parametrize("words,top", [(1, 0), (4, 0), (4, 3), (8, 4), (8, 7), (12, 8), (20, 9), (44, 42), (44, 43), (48, 44)])
def test_fp4_k_extent_follows_the_highest_used_word(gpu, words, top):
    """The fp4 kernel contracts over K blocks of 256 clades and issues only the K = 64 instructions up to the highest
    non-zero word of the gathered rows (found on the device by the gather, read on the device by the kernel): stored
    width `words`, bits only in words <= top, in particular at both ends of a K block (rf_f4.cu)."""
    rng = np.random.default_rng(100 * words + top)
    C, n = 2, 300
    bits = []
    for c in range(C):
        b = np.zeros((n, words), dtype=np.uint64)
        b[:, :top + 1] = rng.integers(0, 2 ** 63, size=(n, top + 1), dtype=np.uint64) & rng.integers(0, 2 ** 63, size=(n, top + 1), dtype=np.uint64)
        b[:, top] |= np.uint64(1) << np.uint64(63 if c else 0)   # the top word is in use, at its first / last bit
        bits.append(b)
    want = np.zeros((C, n, C), dtype=np.int64)
    for a in range(C):
        for c in range(C):
            d = np.bitwise_count(bits[a][:, None, :] ^ bits[c][None, :, :]).sum(-1, dtype=np.int64)
            want[a, :, c] = ((d + 1) ** 2).sum(axis=1)
    with gpu.GpuContext(C) as ctx:
        for c in range(C):
            ctx.trees_set(c, bits[c])
        assert np.array_equal(ctx.psrf_sums([0, 0], 0, n, 1, gpu.ASM_RF_FP4), want)
        assert ctx.last_layout()["padded_bits"] == (top + 1) * 64
        # a shorter evaluation whose rows stop using the top word: the extent shrinks with the gathered rows
        if top > 0:
            for c in range(C):
                bits[c][:100, top] = 0
                ctx.trees_set(c, bits[c])
            got = ctx.psrf_sums([0, 0], 0, 100, 1, gpu.ASM_RF_FP4)
            assert np.array_equal(got, ctx.psrf_sums([0, 0], 0, 100, 1, gpu.ASM_RF_POPC))
            ctx.psrf_sums([0, 0], 0, 100, 1, gpu.ASM_RF_FP4)
            assert ctx.last_layout()["padded_bits"] <= top * 64


@pytest.mark.gpu
def test_trees_from_clade_ids_equal_trees_from_bit_rows(gpu, multi_chains, ragged_chains):
    """asm_trees_set_ids / asm_trees_append_ids build on the GPU the rows asm_encode_bits builds on the host: same S, same
    means, same clade counts; a dictionary that grows between appends widens the stored rows; ASM_NO_CLADE slots and ids
    outside the dictionary behave as documented."""
    for ch in (multi_chains, ragged_chains):
        C, D = ch.n_chains, ch.enc.num_clades
        end = min(ch.n_trees)
        with gpu.GpuContext(C) as a, gpu.GpuContext(C) as b:
            ch.upload(a)
            for c in range(C):
                half = ch.n_trees[c] // 2
                first = ch.ids[c][:half]
                b.trees_set_ids(c, first, int(first.max()) + 1)          # the dictionary as it stood after these trees
                b.trees_append_ids(c, ch.ids[c][half:], D)               # ... and now: wider rows
                assert b.trees_count(c)[0] == ch.n_trees[c]
            for m in (gpu.ASM_RF_POPC, gpu.ASM_RF_I8, gpu.ASM_RF_FP4):
                assert np.array_equal(a.psrf_sums([0] * C, 0, end, 1, m), b.psrf_sums([0] * C, 0, end, 1, m))
            ma, mb = a.psrf_eval([3] * C, 0, end, 1, gpu.ASM_RF_FP4), b.psrf_eval([3] * C, 0, end, 1, gpu.ASM_RF_FP4)
            assert bits_equal(ma, mb)
            for c in range(C):
                assert np.array_equal(a.clade_counts(c, 0, ch.n_trees[c])[:D], b.clade_counts(c, 0, ch.n_trees[c])[:D])
    # unused slots (a multifurcating tree has fewer clades) and a row of its own kind
    ids = np.array([[0, 5, 0xFFFF, 70], [0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF], [129, 128, 127, 126]], dtype=np.uint16)
    want = np.zeros((3, 4), dtype=np.uint64)
    for r, row in enumerate(ids):
        for k in row:
            if k != 0xFFFF:
                want[r, k >> 6] |= np.uint64(1) << np.uint64(k & 63)
    with gpu.GpuContext(2) as ctx:
        ctx.trees_set_ids(0, ids, 130)
        ctx.trees_set(1, want)
        assert ctx.trees_count(0) == (3, 4)
        d = ctx.rf_block(0, 0, 3, 1, 0, 3)
        assert np.array_equal(np.diag(d), [0, 0, 0]) and d[0, 1] == 3 and d[0, 2] == 7
        ctx.trees_set_ids(0, ids, 100)                                   # ids 126..129 are outside a 100-clade dictionary:
        with pytest.raises(gpu.AsmGpuError) as e:                        # reported by the next call that looks at the trees
            ctx.psrf_eval([0, 0], 0, 3, 1, gpu.ASM_RF_POPC)
        assert e.value.code == gpu.ASM_E_RANGE
        ctx.trees_set_ids(0, ids, 130)                                   # a good upload clears the condition
        assert np.isfinite(ctx.psrf_eval([0, 0], 0, 3, 1, gpu.ASM_RF_POPC)).all()
        ctx.trees_set_ids(0, ids, 100)
        with pytest.raises(gpu.AsmGpuError) as e:
            ctx.rf_block(0, 0, 3, 1, 0, 3)
        assert e.value.code == gpu.ASM_E_RANGE
        ctx.trees_set_ids(0, ids, 130)
        with pytest.raises(gpu.AsmGpuError):
            ctx.trees_set_ids(0, ids, 70000)                             # wider than uint16 ids can name
