"""Tuning probe: staged ESS on cfg3 under different lags-per-thread choices per stage (ASM_ESS_R)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
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
res = {}
ref = None
settings = sys.argv[1:] or ["0,0,0,0,0", "4,2,8,4,1", "4,4,8,4,1", "4,2,8,4,2", "4,2,8,2,1", "2,2,8,4,1", "4,2,4,4,1"]
for setting in settings:
    os.environ["ASM_ESS_R"] = setting
    ctx.ess_set_adaptive(True)
    act, ess = ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n)
    if ref is None:
        ref = ess.copy()
    same = bool(np.array_equal(ref.view(np.uint64), ess.view(np.uint64)))
    ctx.timer_mark(0); ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n); ctx.timer_mark(1)
    plain = ctx.timer_elapsed_ms(0, 1)
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n)
    k, kn = ctx.profile_get(L.K_ESS_LAG)
    ctx.profile_enable(False)
    res[setting] = {"step_ms": plain, "lag_ms_profiled": k, "launches": kn, "same_bits": same}
    print(setting, res[setting], flush=True)
print(json.dumps(res))
