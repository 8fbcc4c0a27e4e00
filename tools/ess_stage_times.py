"""One staged + one all-lags ESS evaluation of cfg3 (for an ncu launch list; tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asm_b200 as A
from asm_b200 import _lib as L
import synthgen as G
import bench
from concurrent.futures import ThreadPoolExecutor
n_tr, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000), 1_000_000
ctx = A.GpuContext(1)
host = np.empty((n_tr, n))
def one(j):
    G.synth_ar1(n, bench.MUS[j % 3], bench.PHIS[j % 4], 777 + j, out=host[j])
with ThreadPoolExecutor(32) as ex:
    list(ex.map(one, range(n_tr)))
dev = torch.from_numpy(host).cuda()
for adaptive in (True, False):
    ctx.ess_set_adaptive(adaptive)
    ctx.timer_mark(0); act, ess = ctx.ess_batch_device(dev.data_ptr(), n_tr, n, n); ctx.timer_mark(1)
    print("adaptive", adaptive, "ms", ctx.timer_elapsed_ms(0, 1), "chains", ctx.ess_last_lag_count(), "min ess", float(np.nanmin(ess)))
