"""Host encoder throughput (trees/s) against the thread count, at a cfg5-like tree size.

  python tools/encoder_throughput.py [--taxa 500] [--trees 40000] [--threads 1,2,4,8,16]

Prints one JSON line; the ids of every threaded run are checked against the single-threaded run."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asm_b200 as A  # noqa: E402
import synthgen as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=500)
    ap.add_argument("--trees", type=int, default=40000)
    ap.add_argument("--threads", default="1,2,4,8,16")
    a = ap.parse_args()
    l, r, root, _ = G.synth_tree_chain(a.taxa, a.trees, 1, 20240001, 2, 3.3)
    out = {"taxa": a.taxa, "trees": a.trees, "host_cores": os.cpu_count(), "runs": []}
    ref = None
    for th in [int(x) for x in a.threads.split(",")]:
        os.environ["ASM_ENCODER_THREADS"] = str(th)
        enc = A.Encoder(a.taxa)
        t0 = time.perf_counter()
        ids = enc.add_topologies(l, r, root)
        t1 = time.perf_counter()
        bits = enc.bits(ids)
        t2 = time.perf_counter()
        if ref is None:
            ref = ids
        out["runs"].append({"threads": th, "encode_trees_per_s": a.trees / (t1 - t0), "bits_trees_per_s": a.trees / (t2 - t1),
                            "clades": enc.num_clades, "ids_identical_to_first_run": bool(np.array_equal(ref, ids))})
    base = out["runs"][0]["encode_trees_per_s"]
    for run in out["runs"]:
        run["speedup"] = run["encode_trees_per_s"] / base
    print(json.dumps(out))


if __name__ == "__main__":
    main()
