"""Pair-walk ordering probe for the int8 kernel: kernel time vs supertile edge G (0 = column-major)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asm_b200 as A
from asm_b200 import _lib as L
import bench

res = {}
for name, cfg in {"cfg2": bench.CFG2, "mid_8x12500x500": dict(bench.CFG5, n_chains=8, n_trees=12500),
                  "big_16x25000x500": dict(bench.CFG5, n_chains=16, n_trees=25000)}.items():
    enc, bits, _ = bench.gen_psrf_input(cfg, cfg["n_trees"])
    C, N = cfg["n_chains"], cfg["n_trees"]
    ctx = A.GpuContext(C)
    for c in range(C):
        ctx.trees_set(c, bits[c])
    out = {"D": enc.num_clades, "T": C * N}
    for cta in ("2", "1"):
        if cta == "1":
            os.environ["ASM_I8_1CTA"] = "1"
        for G in ("0", "2", "4", "8", "16", "32", "auto"):
            if cta == "1" and G not in ("0", "auto"):
                continue
            if G == "auto":
                os.environ.pop("ASM_I8_SUPERTILE", None)
            else:
                os.environ["ASM_I8_SUPERTILE"] = G
            ctx.psrf_eval([0] * C, 0, N, 1, L.ASM_RF_I8)
            ctx.profile_enable(True); ctx.profile_reset()
            for _ in range(3):
                ctx.flush_l2(); ctx.psrf_eval([0] * C, 0, N, 1, L.ASM_RF_I8)
            ms, k = ctx.profile_get(L.K_RF_I8)
            ctx.profile_enable(False)
            out[f"cta{cta}_G{G}_ms"] = round(ms / k, 4)
        break  # the 1-CTA flag is read once per process; keep the 2-CTA sweep only
    res[name] = out
    ctx.close()
print(json.dumps(res, indent=1))
