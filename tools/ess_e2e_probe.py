"""Tuning probe: where the end-to-end time of asm_ess_batch (host traces) goes: raw H2D vs compute."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asm_b200 as A
from asm_b200 import _lib as L
import synthgen as G
import bench
from concurrent.futures import ThreadPoolExecutor
n_tr, n = 1000, 1_000_000
ctx = A.GpuContext(1)
host = np.empty((n_tr, n))
def one(j):
    G.synth_ar1(n, bench.MUS[j % 3], bench.PHIS[j % 4], 777 + j, out=host[j])
with ThreadPoolExecutor(32) as ex:
    list(ex.map(one, range(n_tr)))
t, hp = bench.pinned_like(host)
print("pinned", t.is_pinned(), flush=True)
dev = torch.empty((n_tr, n), dtype=torch.float64, device="cuda")
res = {}
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(t, non_blocking=True); torch.cuda.synchronize()
    res[f"h2d_ms_{k}"] = (time.perf_counter() - t0) * 1e3
res["h2d_gbs"] = 8.0 / (res["h2d_ms_2"] * 1e-3)
for k in range(3):
    t0 = time.perf_counter(); act, ess = ctx.ess_batch(hp); res[f"e2e_ms_{k}"] = (time.perf_counter() - t0) * 1e3
for env in sys.argv[1:]:
    os.environ["ASM_ESS_H2D_GROUPS"] = env
    for k in range(2):
        t0 = time.perf_counter(); act, ess = ctx.ess_batch(hp); res[f"e2e_ms_groups{env}_{k}"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(res, indent=1))
