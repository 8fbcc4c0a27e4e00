"""cfg5 (BASELINE.json configs[4]: 32 chains x 50,000 trees x 500 taxa) on one B200: timing of the full
PSRF evaluation and parity evidence at full size.

  * S rows of a few sampled trees are recomputed on the host with numpy XOR + popcount over ALL 1.6M
    column trees (independent of both CUDA kernels) and compared with the GPU table (exact int64);
  * the popcount kernel and the int8 tcgen05 kernel must produce the same full S table (optional, slow);
  * shard partials (asm_shard {i, n}) must add up to the full table.
usage: python tools/cfg5_check.py [--popc] [--chains 32] [--trees 50000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asm_b200 as A  # noqa: E402
import bench  # noqa: E402
from asm_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--popc", action="store_true")
ap.add_argument("--chains", type=int, default=32)
ap.add_argument("--trees", type=int, default=50000)
ap.add_argument("--rows", type=int, default=6)
ap.add_argument("--method", default="i8", choices=["i8", "fp4"])
args = ap.parse_args()

cfg = dict(bench.CFG5, n_chains=args.chains, n_trees=args.trees)
C, N = cfg["n_chains"], cfg["n_trees"]
t0 = time.time()
enc, bits, _ = bench.gen_psrf_input(cfg, N)
out = {"config": f"{C} chains x {N} trees x {cfg['n_taxa']} taxa", "D": enc.num_clades, "words": enc.words,
       "gen_s": round(time.time() - t0, 1)}
ctx = A.GpuContext(C)
for c in range(C):
    ctx.trees_set(c, bits[c])
burnin = [0] * C
METHOD = L.ASM_RF_FP4 if args.method == "fp4" else L.ASM_RF_I8
t0 = time.time()
S = ctx.psrf_sums(burnin, 0, N, 1, METHOD)
out["first_eval_s"] = round(time.time() - t0, 2)
ctx.timer_mark(0)
mean = ctx.psrf_eval(burnin, 0, N, 1, METHOD)
ctx.timer_mark(1)
ms = ctx.timer_elapsed_ms(0, 1)
T = C * N
out.update({"method": args.method, "step_ms": ms, "tree_pairs_per_s": T * T / (ms * 1e-3), "layout": ctx.last_layout(),
            "algorithmic_tflops": enc.num_clades * T * (T - 1) / (ms * 1e-3) / 1e12,
            "psrf_mean_range": [float(np.nanmin(mean[~np.eye(C, dtype=bool)])), float(np.nanmax(mean[~np.eye(C, dtype=bool)]))]})

# host recomputation of sampled rows over all columns
rng = np.random.default_rng(1)
ok = True
t0 = time.time()
for _ in range(args.rows):
    a, r = int(rng.integers(C)), int(rng.integers(N))
    row = bits[a][r]
    for c in range(C):
        d = np.bitwise_count(np.bitwise_xor(bits[c], row[None, :])).sum(axis=1, dtype=np.int64)
        want = int(((d + 1) ** 2).sum())
        if want != int(S[a, r, c]):
            ok = False
            out.setdefault("mismatch", []).append([a, r, c, want, int(S[a, r, c])])
out["host_rows_checked"] = args.rows
out["host_rows_match"] = ok
out["host_check_s"] = round(time.time() - t0, 1)

parts = [ctx.psrf_sums(burnin, 0, N, 1, METHOD, shard=(i, 4)) for i in range(4)]
out["shards_add_up"] = bool(np.array_equal(sum(parts), S))
del parts
if args.method == "fp4":  # the two tensor-core kernels must agree on the full table
    t0 = time.time()
    S1 = ctx.psrf_sums(burnin, 0, N, 1, L.ASM_RF_I8)
    out["i8_eval_s"] = round(time.time() - t0, 2)
    out["fp4_equals_i8"] = bool(np.array_equal(S, S1))
    del S1
if args.popc:
    t0 = time.time()
    S2 = ctx.psrf_sums(burnin, 0, N, 1, L.ASM_RF_POPC)
    out["popc_eval_s"] = round(time.time() - t0, 2)
    out["popc_equals_i8"] = bool(np.array_equal(S, S2))
print(json.dumps(out))
