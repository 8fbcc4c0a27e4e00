"""TreeConvergenceCheck on the GPU -- the offline tool of the reference (src/asm/tools/TreeConvergenceCheck.java):
load two or more BEAST `.trees` files and run TreePSRF.converged once on every unordered pair with a fixed
percentage burn-in (:44-80).  Inputs keep the reference's names and defaults (:21-27):

    python -m asm_b200.tools.tree_convergence_check -tree a.trees b.trees [c.trees ...]
           [-burnin 10] [-smoothing 0.6] [-b 0.05] [-twoSided true] [-checkESS false] [-targetESS 100]

Differences, both forced by the reference itself: the tree distance is Robinson-Foulds on clade bit-vectors
instead of RNNI, and cacheLimit is lifted to the number of trees (the reference leaves it at 1024 and its
distance cache then fails on longer files).
"""
from __future__ import annotations

import argparse
import sys

from asm_b200 import criteria as K


def load_trees(ti: K.TraceInfo, chain: int, path: str) -> int:
    """FastTreeSet stand-in (:94-112): every `tree ...` line of a NEXUS file, taxon numbering from its Translate block."""
    n = 0
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.lstrip().lower().startswith("tree ") and not line.startswith("tree STATE"):
                line = "tree STATE" + line.lstrip()[4:]  # readTreeLogLine only accepts `tree STATE...` (AutoStopMCMC.java:522)
            if ti.readTreeLogLine(chain, line):
                n += 1
    return n


def check_pair(file1: str, file2: str, burnin_pct=10, smoothing=0.6, b=0.05, twoSided=True, checkESS=False, targetESS=100,
               device=0):
    """One iteration of the pair loop (:52-77). Returns (converged, grValues, end)."""
    ti = K.TraceInfo(2, device)                                              # :56-60
    n1, n2 = load_trees(ti, 0, file1), load_trees(ti, 1, file2)
    end = min(n1, n2)                                                        # :69
    crit = K.TreePSRF(smoothing=smoothing, b=b, twoSided=twoSided, checkESS=checkESS, targetESS=targetESS,
                      cacheLimit=max(1024, end + 1, targetESS + 1))          # :61-67
    crit.setup(2, ti)                                                        # :68
    start = burnin_pct * end // 100                                          # :70
    conv = crit.converged([start, start], end)                               # :73
    return conv, crit.grValues, end


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Check convergence of tree files (GPU TreePSRF)", prefix_chars="-")
    ap.add_argument("-tree", dest="trees", nargs="+", required=True)
    ap.add_argument("-burnin", type=int, default=10)
    ap.add_argument("-smoothing", type=float, default=0.6)
    ap.add_argument("-b", type=float, default=0.05)
    ap.add_argument("-twoSided", type=lambda s: s.lower() != "false", default=True)
    ap.add_argument("-checkESS", type=lambda s: s.lower() != "false", default=False)
    ap.add_argument("-targetESS", type=int, default=100)
    a = ap.parse_args(argv)
    if len(a.trees) < 2:
        print("At least two trees should be specified", file=sys.stderr)     # :47-50
        return 1
    for i in range(len(a.trees) - 1):
        for j in range(i + 1, len(a.trees)):
            conv, gr, end = check_pair(a.trees[i], a.trees[j], a.burnin, a.smoothing, a.b, a.twoSided, a.checkESS, a.targetESS)
            print(f"Comparing {a.trees[i]} and {a.trees[j]} ({end} trees): psrf1mean = {gr[0]:.3f} psrf2mean = {gr[1]:.3f}")
            print(("Converged" if conv else "Not converged") + " according to the GRLike criterion")   # :75
    print("Done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
