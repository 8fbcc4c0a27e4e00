"""Tuning probe: host-side timeline of asm_ess_batch on cfg3 (ASM_ESS_TRACE)."""
import os, sys, time
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
ctx.ess_batch(hp); ctx.ess_batch(hp)
os.environ["ASM_ESS_TRACE"] = "1"
t0 = time.perf_counter(); ctx.ess_batch(hp); print("e2e ms", (time.perf_counter() - t0) * 1e3)
