"""Tuning probe: staged ESS on cfg3 under different stage bounds (ASM_ESS_BOUNDS)."""
import json, os, sys
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
dev = torch.from_numpy(host).cuda()
ref = None
for b in ["default"] + sys.argv[1:]:
    if b != "default":
        os.environ["ASM_ESS_BOUNDS"] = b
    act, ess = ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n)
    if ref is None: ref = ess.copy()
    ts = []
    for k in range(3):
        ctx.timer_mark(0); ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n); ctx.timer_mark(1); ts.append(round(ctx.timer_elapsed_ms(0, 1), 2))
    print(b, ts, "chains", ctx.ess_last_lag_count(), "same", bool(np.array_equal(ref.view(np.uint64), ess.view(np.uint64))), flush=True)
