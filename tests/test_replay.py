"""Stop-decision identity over a whole run: synthetic BEAST-format logs shaped like
examples/dna.xml and examples/yang.xml are replayed check by check, exactly as LogWatcherThread feeds the
criteria (AutoStopMCMC.java:355-426), into the oracle and into the GPU criteria."""
import numpy as np
import pytest

import oracle as O
import replay
from conftest import bits_equal

SCENARIOS = {
    # examples/dna.xml:39-41 -- 10 taxa, GRLike cacheLimit=200 sampleSize=10 twoSided smoothing=0.9 targetESS=100,
    # TraceESS targetESS=100, Automatic burn-in
    "dna": dict(n_taxa=10, beta=1.2, n_samples=420, psrf=dict(b=0.05, cacheLimit=200, sampleSize=10, twoSided=True,
                                                               smoothing=0.9, targetESS=100),
                ess_target=100, strategy="Automatic", percent=10),
    # examples/yang.xml:65-67 -- 35 taxa, burnInStrat=Running burnInPercent=5, GRLike sampleSize=1 targetESS=2
    # cacheLimit=128 (delta doubles three times on the way); b and the TraceESS target are relaxed so that the synthetic
    # chains do reach a stop decision within 600 checks
    "yang": dict(n_taxa=35, beta=2.0, n_samples=600, psrf=dict(b=0.3, sampleSize=1, targetESS=2, cacheLimit=128),
                 ess_target=20, strategy="Running", percent=5),
}


def _run(tmp_path, sc, gpu_mod=None, method=1):
    prefixes = replay.make_run(tmp_path, sc["n_taxa"], 2, sc["n_samples"], sc["beta"], seed=5)
    gpu_criteria = None
    if gpu_mod is not None:
        from asm_b200 import criteria as K
        gpu_criteria = [K.TreePSRF(method=method, **sc["psrf"]), K.TraceESS(targetESS=sc["ess_target"])]
    rp = replay.Replay(prefixes, sc["n_taxa"], gpu_criteria, O.TreePSRF(**sc["psrf"]), sc["ess_target"],
                       burnin_strategy=sc["strategy"], burnin_percent=sc["percent"])
    while rp.step() is not None:
        pass
    rp.close()
    return rp


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_replay_oracle_only(tmp_path, name):
    """CPU: the harness reproduces the watcher's bookkeeping and the oracle criteria reach a decision."""
    rp = _run(tmp_path, SCENARIOS[name])
    h = rp.history
    assert len(h) >= SCENARIOS[name]["n_samples"]
    # first round: the header line costs the trace logs a pass, so trees run one ahead (AutoStopMCMC.java:361-375)
    assert (h[0]["n_logs"], h[0]["n_trees"]) == (1, 2) and (h[100]["n_logs"], h[100]["n_trees"]) == (101, 102)
    assert h[0]["end"] == 0 and np.isnan(h[0]["psrf_oracle"][1][0]) and not h[0]["converged_oracle"]
    first = next((r["end"] for r in h if r["converged_oracle"]), None)
    assert first is not None and first > 50
    if name == "yang":
        assert rp.oracle_psrf.delta == 8  # end/delta > cacheLimit at 129, 258, 516 (TreePSRF.java:120-124)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_replay_oracle_matches_frozen_expectations(tmp_path, name):
    """CPU: the oracle's per-check numbers for the replayed runs are frozen in tests/golden/replay_<name>_expected.json
    (tests/golden/make_replay_expected.py); a change of the oracle, the generators or the harness shows up here.
    The same files are what a maintainer compares the Java's output with."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    want = json.load(open(os.path.join(here, "golden", f"replay_{name}_expected.json")))["checks"]
    rp = _run(tmp_path, SCENARIOS[name])
    assert len(rp.history) == len(want)
    hexbits = lambda x: [format(int(v), "016x") for v in np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64).view(np.uint64)]
    for r, w in zip(rp.history, want):
        po, eo = r["psrf_oracle"], r["ess_oracle"]
        assert r["end"] == w["end"] and r["burnin"] == w["burnin"]
        assert [bool(po[0]), hexbits(po[1]), int(po[2]), int(po[3])] == w["psrf"], f"TreePSRF at check {r['end']}"
        assert [bool(eo[0]), hexbits(eo[1]), hexbits(eo[2])[0]] == w["ess"], f"TraceESS at check {r['end']}"
        assert bool(r["converged_oracle"]) == w["converged"]


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_replay_stop_decision_identical(gpu, tmp_path, name, method):
    rp = _run(tmp_path, SCENARIOS[name], gpu_mod=gpu, method=method)
    h = rp.history
    assert len(h) >= SCENARIOS[name]["n_samples"]
    for r in h:
        assert r["burnin"] == r["burnin_gpu"], f"BurnInDetector differs at check {r['end']}"
        po, pg = r["psrf_oracle"], r["psrf_gpu"]
        assert po[0] == pg[0], f"TreePSRF decision differs at check {r['end']}"
        assert bits_equal(po[1], pg[1]), f"grValues differ at check {r['end']}: {po[1]} vs {pg[1]}"
        assert po[2:] == pg[2:], f"delta/start0 differ at check {r['end']}"
        eo, eg = r["ess_oracle"], r["ess_gpu"]
        assert eo[0] == eg[0] and bits_equal(eo[1], eg[1]) and bits_equal(np.array([eo[2]]), np.array([eg[2]]))
        assert r["converged_oracle"] == r["converged_gpu"]
    first_o = next((r["end"] for r in h if r["converged_oracle"]), None)
    first_g = next((r["end"] for r in h if r["converged_gpu"]), None)
    assert first_o == first_g and first_o is not None


@pytest.mark.gpu
def test_replay_more_than_two_chains(gpu, tmp_path):
    """cfg1 of BASELINE.json names 4 chains: beyond what the reference TreePSRF accepts (TreePSRF.java:311-313);
    the generalisation (every ordered pair) is compared with the oracle's."""
    from asm_b200 import criteria as K
    prefixes = replay.make_run(tmp_path, 12, 4, 160, 1.4, seed=9)
    kw = dict(b=0.08, cacheLimit=64, sampleSize=5, smoothing=0.9, targetESS=50)
    rp = replay.Replay(prefixes, 12, [K.TreePSRF(**kw), K.TraceESS(targetESS=10)], O.TreePSRF(**kw), 10,
                       burnin_strategy="Running", burnin_percent=10)
    while rp.step() is not None:
        pass
    for r in rp.history:
        assert r["psrf_oracle"][0] == r["psrf_gpu"][0] and r["psrf_oracle"][2:] == r["psrf_gpu"][2:]
        go, gg = r["psrf_oracle"][1], r["psrf_gpu"][1]
        off = ~np.eye(4, dtype=bool)
        assert bits_equal(go[off], gg[off])
        assert bits_equal(r["ess_oracle"][1], r["ess_gpu"][1]) and r["converged_oracle"] == r["converged_gpu"]
    assert rp.oracle_psrf.delta == 4


@pytest.mark.gpu
def test_tree_ess_pseudo_ess_and_one_sided(gpu, small_chains):
    """TreeESS.pseudoESS (TreeESS.java:117-166) through asm_rf_block + asm_ess_batch, reached via the one-sided
    TreePSRF with checkESS (TreePSRF.java:232-247)."""
    from asm_b200 import criteria as K
    ch = small_chains
    ti = K.TraceInfo(2)
    ti.setTaxonCount(ch.n_taxa)
    for c in range(2):
        l, r, root = ch.topo[c]
        for i in range(300):
            ti.addTree(c, l[i], r[i], int(root[i]))
    kw = dict(b=10.0, twoSided=False, checkESS=True, targetESS=20, sampleSize=5, cacheLimit=1024, smoothing=1.0)
    g, o = K.TreePSRF(**kw), O.TreePSRF(**kw)
    g.setup(2, ti)
    o.setup(2, ch.ts)
    for end in (10, 21, 60, 150, 300):
        assert g.converged([0, 0], end) == o.converged([0, 0], end), end
        assert bits_equal(g.grValues[:1], o.grValues[:1])
    assert g.converged([0, 0], 301) is False and g.threw  # past the stored trees: caught, false (TreePSRF.java:252-258)


@pytest.mark.gpu
def test_tree_convergence_check_tool(gpu, tmp_path):
    """The offline batch caller (tools/TreeConvergenceCheck.java:44-80): one converged() call per pair of files."""
    from asm_b200.tools import tree_convergence_check as T
    prefixes = replay.make_run(tmp_path, 14, 3, 300, 1.6, seed=21)
    files = [p + ".trees" for p in prefixes]
    # oracle: same trees, same single call
    import oracle as O2
    for (i, j) in [(0, 1), (0, 2), (1, 2)]:
        conv, gr, end = T.check_pair(files[i], files[j], burnin_pct=10, smoothing=0.9, b=0.2, targetESS=200)
        ts = O2.TreeSet(2, 14)
        for c, f in enumerate((files[i], files[j])):
            for line in open(f):
                if line.startswith("tree STATE"):
                    l, r, root = replay.parse_newick(line[line.index("("):], 14)
                    ts.add(c, l[None, :], r[None, :], np.array([root], dtype=np.int32))
        crit = O2.TreePSRF(smoothing=0.9, b=0.2, targetESS=200, cacheLimit=max(1024, end + 1))
        crit.setup(2, ts)
        start = 10 * end // 100
        assert crit.converged([start, start], end) == conv and end == 302
        assert bits_equal(crit.grValues, gr)
    assert T.main(["-tree", files[0], files[1], "-b", "0.2"]) == 0
