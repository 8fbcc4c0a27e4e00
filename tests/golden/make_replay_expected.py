"""Freezes what the ORACLE says at every check of the two replayed runs (tests/test_replay.py SCENARIOS), and can
leave the BEAST-format log files behind for a maintainer who wants to pin the oracle against the real Java.

    python tests/golden/make_replay_expected.py                 # rewrites tests/golden/replay_<name>_expected.json
    python tests/golden/make_replay_expected.py --keep-logs DIR # also writes DIR/<name>/chain<i>-run.{log,trees}

The JSON holds, per check: end (m_nLastReported), burnin[], TreePSRF (decision, grValues, delta, start0) and TraceESS
(decision, currentESSs, minESS); doubles as 16-digit hex of their bit patterns.  PARITY UNPINNED: these are the
oracle's numbers, not the Java's.  INTEGRATION.md section 5 says how to obtain the Java's for the same files
(asm.tools.ASMEmulator replays pre-recorded logs, ASMEmulator.java:16-41).
"""
import argparse
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)

import oracle as O  # noqa: E402
import replay  # noqa: E402
from test_replay import SCENARIOS  # noqa: E402


def hexbits(x):
    return [format(int(v), "016x") for v in np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64).view(np.uint64)]


def run_oracle(name, workdir):
    sc = SCENARIOS[name]
    prefixes = replay.make_run(workdir, sc["n_taxa"], 2, sc["n_samples"], sc["beta"], seed=5)
    rp = replay.Replay(prefixes, sc["n_taxa"], None, O.TreePSRF(**sc["psrf"]), sc["ess_target"],
                       burnin_strategy=sc["strategy"], burnin_percent=sc["percent"])
    while rp.step() is not None:
        pass
    rp.close()
    out = []
    for r in rp.history:
        po, eo = r["psrf_oracle"], r["ess_oracle"]
        out.append({"end": r["end"], "burnin": r["burnin"], "psrf": [bool(po[0]), hexbits(po[1]), int(po[2]), int(po[3])],
                    "ess": [bool(eo[0]), hexbits(eo[1]), hexbits(eo[2])[0]], "converged": bool(r["converged_oracle"])})
    return {"scenario": {k: v for k, v in sc.items()}, "seed": 5, "checks": out}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--keep-logs", default=None)
    args = ap.parse_args()
    for name in SCENARIOS:
        tmp = tempfile.mkdtemp(prefix="asm_replay_")
        try:
            data = run_oracle(name, tmp)
            with open(os.path.join(HERE, f"replay_{name}_expected.json"), "w") as f:
                json.dump(data, f, separators=(",", ":"))
            if args.keep_logs:
                dst = os.path.join(args.keep_logs, name)
                os.makedirs(dst, exist_ok=True)
                for fn in os.listdir(tmp):
                    shutil.copy(os.path.join(tmp, fn), os.path.join(dst, fn))
            first = next((c["end"] for c in data["checks"] if c["converged"]), None)
            print(name, len(data["checks"]), "checks, first stop at", first)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
