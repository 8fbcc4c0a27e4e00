"""cfg5-shaped probe: full cfg5 dictionary (D ~ 10.8k), first `end` trees of every chain; kernel time vs G."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asm_b200 as A
from asm_b200 import _lib as L
import bench
end = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
Gs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2", "4", "8", "11", "16"]
cfg = bench.CFG5
enc, bits, _ = bench.gen_psrf_input(cfg, cfg["n_trees"])
C = cfg["n_chains"]
ctx = A.GpuContext(C)
for c in range(C):
    ctx.trees_set(c, bits[c][:end])
out = {"D": enc.num_clades, "T": C * end}
for G in Gs:
    os.environ["ASM_I8_SUPERTILE"] = G
    ctx.psrf_eval([0] * C, 0, end, 1, L.ASM_RF_I8)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(2):
        ctx.flush_l2(); ctx.psrf_eval([0] * C, 0, end, 1, L.ASM_RF_I8)
    ms, k = ctx.profile_get(L.K_RF_I8)
    ctx.profile_enable(False)
    out[f"G{G}_ms"] = round(ms / k, 3)
print(json.dumps(out))
