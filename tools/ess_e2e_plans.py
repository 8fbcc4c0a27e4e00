"""Tuning probe: end-to-end asm_ess_batch on cfg3 under different upload chunk plans (ASM_ESS_H2D_PLAN)."""
import os, sys, time, json
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
ref = ctx.ess_batch(hp)[1]
res = {}
for plan in ["default"] + sys.argv[1:]:
    if plan != "default":
        os.environ["ASM_ESS_H2D_PLAN"] = plan
    ts = []
    for k in range(3):
        t0 = time.perf_counter(); act, ess = ctx.ess_batch(hp); ts.append((time.perf_counter() - t0) * 1e3)
    res[plan] = {"ms": [round(x, 1) for x in ts], "same": bool(np.array_equal(ref.view(np.uint64), ess.view(np.uint64)))}
    print(plan, res[plan], flush=True)
