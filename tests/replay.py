"""Replay harness (test infrastructure), modelled on the reference's ASMEmulator
(src/asm/tools/ASMEmulator.java:27-39, EmulatedMCMCChain.java:79-104): pre-recorded BEAST log and tree
files are fed line by line to the stopping criteria the way AutoStopMCMC.LogWatcherThread does
(AutoStopMCMC.java:355-426), once into the CPU oracle and once into the GPU criteria, and every check's
decision and logged numbers are compared.

The logs here are SYNTHETIC files in BEAST's formats (TSV trace log with a `Sample` header, NEXUS tree log
with a Translate block), shaped like examples/dna.xml (10 taxa) and examples/yang.xml (35 taxa); BEAST
itself cannot run in this image, so these are not BEAST runs.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

import oracle as O


# ------------------------------------------------------------------------------------------ file writers
def newick(left, right, root, rng) -> str:
    """Newick with 1-based integer labels (the NEXUS translate index) and branch lengths."""
    def rec(v):
        bl = f":{rng.uniform(0.001, 0.2):.6g}"
        if left[v] < 0:
            return f"{v + 1}{bl}"
        return f"({rec(left[v])},{rec(right[v])}){bl}"
    return "(" + rec(left[root]) + "," + rec(right[root]) + "):0.0;"


def write_chain_files(path_prefix: str, n_taxa: int, topo, traces: np.ndarray, labels: List[str], log_every=1000,
                      seed=0):
    """traces: [n_cols-1][n_samples] (column 0, the sample number, is added here)."""
    rng = np.random.default_rng(seed)
    left, right, root = topo
    n = traces.shape[1]
    with open(path_prefix + ".log", "w") as f:
        f.write("# BEAST v2.7.5 (synthetic replay log)\n#\n# seed " + str(seed) + "\n")
        f.write("Sample\t" + "\t".join(labels) + "\n")
        for i in range(n):
            f.write(str(i * log_every) + "\t" + "\t".join(repr(float(v)) for v in traces[:, i]) + "\t\n")
    with open(path_prefix + ".trees", "w") as f:
        f.write("#NEXUS\n\nBegin taxa;\n\tDimensions ntax=%d;\n\t\tTaxlabels\n" % n_taxa)
        for t in range(n_taxa):
            f.write(f"\t\t\ttaxon{t}\n")
        f.write("\t\t\t;\nEnd;\nBegin trees;\n\tTranslate\n")
        for t in range(n_taxa):
            f.write(f"\t\t{t + 1:4d} taxon{t}" + ("," if t + 1 < n_taxa else "") + "\n")
        f.write(";\n")
        for i in range(left.shape[0]):
            f.write(f"tree STATE_{i * log_every} = {newick(left[i], right[i], int(root[i]), rng)}\n")
        f.write("End;\n")


# ------------------------------------------------------------------------------------------ python-side parsing (oracle feed)
def parse_newick(s: str, n_taxa: int, offset: int = 1):
    """Integer-labelled binary Newick -> (left, right, root) child arrays, for the oracle."""
    nn = 2 * n_taxa - 1
    left = np.full(nn, -1, dtype=np.int32)
    right = np.full(nn, -1, dtype=np.int32)
    stack, marks, nxt, i = [], [], n_taxa, 0
    while i < len(s) and s[i] != ";":
        c = s[i]
        if c == "(":
            marks.append(len(stack))
            i += 1
        elif c == ")":
            m = marks.pop()
            assert len(stack) - m == 2
            r, l = stack.pop(), stack.pop()
            left[nxt], right[nxt] = l, r
            stack.append(nxt)
            nxt += 1
            i += 1
        elif c == "[":
            i = s.index("]", i) + 1
        elif c == ":":
            i += 1
            while i < len(s) and (s[i].isdigit() or s[i] in ".eE-+"):
                i += 1
        elif c in ", \t":
            i += 1
        else:
            j = i
            while s[j].isdigit():
                j += 1
            stack.append(int(s[i:j]) - offset)
            i = j
    assert len(stack) == 1 and nxt == nn
    return left, right, stack[0]


def java_split_ws(s: str):
    import re
    parts = re.split(r"\s+", s)
    while parts and parts[-1] == "":
        parts.pop()
    return parts


class _Files:
    def __init__(self, prefixes):
        self.logs = [open(p + ".log") for p in prefixes]
        self.trees = [open(p + ".trees") for p in prefixes]

    def close(self):
        for f in self.logs + self.trees:
            f.close()


class Replay:
    """One watcher loop driving two sets of criteria in lockstep."""

    def __init__(self, prefixes, n_taxa, gpu_criteria, oracle_psrf: Optional[O.TreePSRF], trace_ess_target: Optional[int],
                 traces="posterior,prior,likelihood", burnin_strategy="Automatic", burnin_percent=10, device=0):
        from asm_b200 import criteria as K
        self.K = K
        self.n_chains = len(prefixes)
        self.n_taxa = n_taxa
        self.files = _Files(prefixes)
        self.gpu_criteria = gpu_criteria or []
        self.ti = K.TraceInfo(self.n_chains, device) if gpu_criteria else None  # oracle-only replay needs no GPU
        self.trace_names = traces
        self.oracle_psrf = oracle_psrf
        self.trace_ess_target = trace_ess_target
        self.strategy, self.percent = burnin_strategy, burnin_percent
        # oracle-side stores (TraceInfo.logLines / trees)
        self.ts = O.TreeSet(self.n_chains, n_taxa)
        self.cols: List[List[List[float]]] = [[] for _ in range(self.n_chains)]
        self.labels = None
        self.map = None
        self.n_last_reported = 0  # m_nLastReported
        self.history = []
        for c in self.gpu_criteria:
            c.setup(self.n_chains, self.ti)
        if oracle_psrf is not None:
            oracle_psrf.setup(self.n_chains, self.ts)

    # readLogLines (AutoStopMCMC.java:430-479)
    def _read_log_lines(self, chain) -> bool:
        f = self.files.logs[chain]
        while True:
            line = f.readline()
            if not line:
                return False
            line = line.rstrip("\n")
            toks = java_split_ws(line)
            if line.startswith("#") or len(toks) == 1:
                continue
            break
        got = self.ti.readLogLine(chain, line) if self.ti else not line.startswith("Sample")  # GPU side
        if line.startswith("Sample"):
            self.labels = toks
            assert not got
            return False
        vals = []
        for t in toks:
            try:
                vals.append(float(t))
            except ValueError:
                vals.append(0.0)
        if not self.cols[chain]:
            self.cols[chain] = [[] for _ in vals]
        for j, v in enumerate(vals):
            self.cols[chain][j].append(v)
        assert got
        return True

    # readTreeLogLine (AutoStopMCMC.java:486-538)
    def _read_tree_log_line(self, chain) -> bool:
        f = self.files.trees[chain]
        while True:
            line = f.readline()
            if not line:
                return False
            line = line.rstrip("\n")
            got = self.ti.readTreeLogLine(chain, line) if self.ti else line.startswith("tree STATE")  # GPU side
            if line.startswith("tree STATE"):
                l, r, root = parse_newick(line[line.index("("):], self.n_taxa)
                self.ts.add(chain, l[None, :], r[None, :], np.array([root], dtype=np.int32))
                assert got
                return True
            assert not got

    def _logs_arrays(self, end):
        # everything read so far ([n_cols][n_read]); the criteria only index [burnin, end)
        return [np.array(self.cols[c], dtype=np.float64) for c in range(self.n_chains)]

    def step(self):
        """One pass of the `while (true)` body (AutoStopMCMC.java:355-390). Returns the record of the check."""
        n2 = 2 * self.n_chains
        n_read = 0
        while n_read < n2:
            done = [False] * n2  # re-created on every pass, as in the Java (:361)
            progressed = False
            for i in range(n2):
                if not done[i]:
                    ok = self._read_log_lines(i // 2) if i % 2 == 0 else self._read_tree_log_line(i // 2)
                    if ok:
                        n_read += 1
                        done[i] = True
                        progressed = True
            if n_read < n2 and not progressed:
                return None  # files exhausted (the Java would sleep and poll)
        rec = self.check()
        self.n_last_reported += 1
        return rec

    def check(self):
        end = self.n_last_reported
        if self.map is None:
            names = [t.strip() for t in self.trace_names.split(",")]
            self.map = [self.labels.index(t) for t in names]
            if self.ti:
                assert self.ti.getMap() == self.map
        logs = self._logs_arrays(end)
        burnin = O.burnin(logs, self.map, self.strategy, self.percent, end) if end > 0 else np.zeros(self.n_chains, np.int32)
        rec = {"end": end, "burnin": burnin.tolist(), "n_trees": self.ts.num_trees(0), "n_logs": len(self.cols[0][0])}
        if self.ti:  # the GPU side gets its burn-in from the product's BurnInDetector (BurnInDetector.java:21-36)
            burnin_gpu = self.K.BurnInDetector(self.ti, self.strategy, self.percent).burnIn(end)
            rec["burnin_gpu"] = burnin_gpu.tolist()
        else:
            burnin_gpu = burnin
        # ---- oracle ----
        conv_o = True
        if self.oracle_psrf is not None:
            r = self.oracle_psrf.converged(burnin, end)
            rec["psrf_oracle"] = (r, self.oracle_psrf.grValues.copy(), self.oracle_psrf.delta, self.oracle_psrf.start0)
            conv_o &= r
        if self.trace_ess_target is not None:
            r, ess, mn = O.trace_ess_converged(logs, burnin, end, self.map, self.trace_ess_target)
            rec["ess_oracle"] = (r, ess, mn)
            conv_o &= r
        # ---- GPU ----
        conv_g = True
        for c in self.gpu_criteria:
            r = c.converged(burnin_gpu, end)
            conv_g &= r
            if isinstance(c, self.K.TreePSRF):
                rec["psrf_gpu"] = (r, c.grValues.copy(), c.delta, c.start0)
            else:
                rec["ess_gpu"] = (r, c.currentESSs, c.minESS)
        rec["converged_oracle"], rec["converged_gpu"] = conv_o, conv_g
        self.history.append(rec)
        return rec

    def close(self):
        self.files.close()


def make_run(tmpdir, n_taxa, n_chains, n_samples, beta, seed, burn_len=40):
    """Synthetic chains -> files on disk.  Traces: AR(1) around a level that relaxes from a bad start."""
    from asm_b200 import _lib as L
    import synthgen as G
    prefixes = []
    for c in range(n_chains):
        l, r, root, _ = G.synth_tree_chain(n_taxa, n_samples + 2, 11, seed + c, 2, beta)
        rng = np.random.default_rng(seed + 100 + c)
        t = np.arange(n_samples + 2)
        cols = []
        for j, (mu, phi) in enumerate([(-5000.0, 0.6), (-4800.0, 0.5), (-200.0, 0.3), (0.5, 0.7)]):
            x = G.synth_ar1(n_samples + 2, 0.0, phi, seed * 7 + c * 13 + j)
            cols.append(mu + x * (1 + j) - 300.0 * np.exp(-t / (burn_len / 3.0)) * (1 + 0.1 * c))
        cols[0] = cols[1] + cols[2]  # posterior = likelihood + prior
        p = os.path.join(str(tmpdir), f"chain{c}-run")
        write_chain_files(p, n_taxa, (l, r, root), np.stack(cols), ["posterior", "likelihood", "prior", "kappa"], seed=seed + c)
        prefixes.append(p)
    return prefixes
