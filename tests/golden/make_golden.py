"""Regenerates tests/golden/*.npz.

The reference ships no golden vectors and cannot be run here (no JVM; BEAST 2 / BEASTLabs absent), so
these fixtures are produced by the CPU oracle (oracle/asm_oracle.c, a line-by-line restatement of the
cited Java) on small seeded inputs, and are committed so that (a) a later change to the oracle that
alters any result is caught, and (b) the GPU box, which has no /root/reference, checks the CUDA path
against values frozen at a known-good commit.  PARITY UNPINNED with respect to the Java itself.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import asm_b200 as A  # noqa: E402  (host encoder + generators only; no GPU involved)
import oracle as O  # noqa: E402
from asm_b200 import _lib as L  # noqa: E402
import synthgen as G


def psrf_fixture():
    n_taxa, n_trees, C = 16, 90, 3
    enc, ts = A.Encoder(n_taxa), O.TreeSet(C, n_taxa)
    topo, bits = [], []
    for c in range(C):
        l, r, root, _ = G.synth_tree_chain(n_taxa, n_trees, 7, 500 + c, 2, 1.8)
        ts.add(c, l, r, root)
        topo.append((l, r, root))
        bits.append(enc.add_topologies(l, r, root))
    words = enc.words
    bits = [enc.bits(b, words) for b in bits]
    out = {"n_taxa": n_taxa, "n_trees": n_trees, "n_chains": C, "words": words, "num_clades": ts.num_clades}
    for c in range(C):
        out[f"left{c}"], out[f"right{c}"], out[f"root{c}"] = topo[c]
        out[f"bits{c}"] = bits[c]
        out[f"clade_counts{c}"] = ts.clade_counts(c, 5, 80)
    out["rf_0_2"] = ts.rf_block(0, 0, n_trees, 2, 0, n_trees)
    cases = [([0, 0, 0], 0, 90, 1), ([9, 20, 3], 9, 88, 1), ([7, 0, 30], 4, 88, 4)]
    out["cases"] = np.array([b + [s, e, d] for b, s, e, d in cases], dtype=np.int32)
    for k, (b, s, e, d) in enumerate(cases):
        out[f"S{k}"] = ts.psrf_sums(b, s, e, d).astype(np.int64)
        means = np.full((C, C), np.nan)
        rows = np.full((C, (e - s + d - 1) // d, C), np.nan)
        for a in range(C):
            for bb in range(C):
                if a != bb:
                    m, r = ts.psrf_rows_mean(a, bb, s, b, e, d, want_rows=True)
                    means[a, bb] = m
                    rows[a, :len(r), bb] = r
        out[f"mean{k}"], out[f"rows{k}"] = means, rows
    np.savez_compressed(os.path.join(HERE, "psrf_small.npz"), **out)


def ess_fixture():
    lens = [1, 2, 7, 64, 1000, 2003, 5000]
    out = {"lens": np.array(lens)}
    for n in lens:
        x = np.stack([G.synth_ar1(n, mu, phi, 900 + i) for i, (mu, phi) in
                      enumerate([(0.0, 0.0), (1.0, 0.5), (-1e4, 0.9), (0.0, 0.99)])])
        act, ess = O.ess_batch(x, literal=(n <= 2003))
        out[f"x{n}"], out[f"act{n}"], out[f"ess{n}"] = x, act, ess
    np.savez_compressed(os.path.join(HERE, "ess_small.npz"), **out)


if __name__ == "__main__":
    psrf_fixture()
    ess_fixture()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
