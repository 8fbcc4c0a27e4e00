"""Timing probe for the staged ESS path: which stage costs what (tuning aid, not a benchmark)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import asm_b200 as A
from asm_b200 import _lib as L
import synthgen as G
import bench

n_tr, n = 1000, 1_000_000
ctx = A.GpuContext(1)
host = np.empty((n_tr, n))
res = {}
for name, phis in {"phi0_only_stage0": [0.0], "phi099_all": [0.99], "mix": bench.PHIS}.items():
    from concurrent.futures import ThreadPoolExecutor
    def one(j):
        G.synth_ar1(n, 0.0, phis[j % len(phis)], 777 + j, out=host[j])
    with ThreadPoolExecutor(32) as ex:
        list(ex.map(one, range(n_tr)))
    dev = torch.from_numpy(host).cuda()
    for adaptive in (True, False):
        ctx.ess_set_adaptive(adaptive)
        ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n)
        ctx.profile_enable(True); ctx.profile_reset()
        ctx.timer_mark(0); ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n); ctx.timer_mark(1)
        k, kn = ctx.profile_get(L.K_ESS_LAG); f, fn = ctx.profile_get(L.K_ESS_FINISH)
        ctx.profile_enable(False)
        res[f"{name}_adaptive{int(adaptive)}"] = {"step_ms_profiled": ctx.timer_elapsed_ms(0, 1), "lag_ms": k, "lag_launches": kn, "scan_ms": f,
                                                  "chains": ctx.ess_last_lag_count()}
    del dev
print(json.dumps(res, indent=1))
