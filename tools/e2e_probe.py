"""Where the end-to-end cfg2 step goes: host-clock time of every C-ABI call of one check (asm_trees_set per chain,
asm_psrf_eval), medians over repeats, plus the per-kernel-family event times of the same step."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench as B
from asm_b200 import _lib as L

cfg = B.CFG2
C, n = cfg["n_chains"], cfg["n_trees"]
enc, bits, _ = B.gen_psrf_input(cfg, n)
pinned = [B.pinned_like(b) for b in bits]
pageable = [np.array(b, copy=True) for b in bits]
ids16 = [B.pinned_like(i.astype(np.uint16)) for i in enc.chain_ids]
D = int(enc.num_clades)
out = {"D": int(enc.num_clades), "words": int(enc.words), "bytes_per_chain": int(bits[0].nbytes),
       "id_bytes_per_chain": int(enc.chain_ids[0].size * 2)}
with L.GpuContext(C, device=0) as ctx:
    for label, src in (("pinned", [p[1] for p in pinned]), ("pageable", pageable), ("pinned_ids", [p[1] for p in ids16])):
        put = (lambda c, a: ctx.trees_set_ids(c, a, D)) if label == "pinned_ids" else ctx.trees_set
        for c in range(C):
            put(c, src[c])
        ctx.psrf_eval([0] * C, 0, n, 1, L.ASM_RF_FP4)
        t_set, t_eval, t_all = [], [], []
        for _ in range(30):
            ctx.flush_l2()
            ctx.sync() if hasattr(ctx, "sync") else None
            t0 = time.perf_counter()
            per = []
            for c in range(C):
                t1 = time.perf_counter()
                put(c, src[c])
                per.append(time.perf_counter() - t1)
            t2 = time.perf_counter()
            ctx.psrf_eval([0] * C, 0, n, 1, L.ASM_RF_FP4)
            t3 = time.perf_counter()
            t_set.append(per)
            t_eval.append(t3 - t2)
            t_all.append(t3 - t0)
        out[label] = {"trees_set_ms_per_chain": [round(1e3 * float(x), 4) for x in np.median(np.array(t_set), axis=0)],
                      "psrf_eval_ms": round(1e3 * float(np.median(t_eval)), 4), "whole_ms": round(1e3 * float(np.median(t_all)), 4)}
    # resident step by host clock, and family event times
    t = []
    for _ in range(30):
        ctx.flush_l2()
        t0 = time.perf_counter()
        ctx.psrf_eval([0] * C, 0, n, 1, L.ASM_RF_FP4)
        t.append(time.perf_counter() - t0)
    out["resident_eval_host_clock_ms"] = round(1e3 * float(np.median(t)), 4)
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(10):
        ctx.flush_l2()
        ctx.psrf_eval([0] * C, 0, n, 1, L.ASM_RF_FP4)
    out["family_ms"] = {k: round(ctx.profile_get(f)[0] / 10, 5) for k, f in
                        (("gather", L.K_GATHER), ("rf", L.K_RF_I8), ("finish", L.K_PSRF_FINISH), ("misc", L.K_MISC))}
    ctx.profile_enable(False)
print(json.dumps(out, indent=1))
