"""Per-check latency of the criteria as the live runner uses them (one converged() per new sample), as a function
of the number of trees per chain: where the GPU path starts to pay off.  2 chains, 35 taxa (yang.xml shape)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asm_b200 as A
import oracle as O
from asm_b200 import _lib as L
import synthgen as G

n_taxa, C = 35, 2
res = []
for m in (100, 200, 1000, 5000, 20000, 50000):
    enc, ts = A.Encoder(n_taxa), O.TreeSet(C, n_taxa)
    bits = []
    for c in range(C):
        l, r, root, _ = G.synth_tree_chain(n_taxa, m, 1, 50 + c, 2, 2.4)
        ids = enc.add_topologies(l, r, root)
        ts.add(c, l, r, root)
        bits.append(ids)
    bits = [enc.bits(b) for b in bits]
    x = np.stack([np.stack([G.synth_ar1(m, 0.0, 0.9, 10 * c + j) for j in range(4)]) for c in range(C)])
    with A.GpuContext(C) as ctx:
        for c in range(C):
            ctx.trees_set(c, bits[c])
            ctx.traces_append(c, np.ascontiguousarray(x[c].T))
        burn, start = [m // 10] * C, m // 10
        row = {"trees_per_chain": m, "D": enc.num_clades}
        for name, meth in (("i8", L.ASM_RF_I8), ("popc", L.ASM_RF_POPC)):
            ctx.psrf_eval(burn, start, m, 1, meth)
            t = []
            for _ in range(10 if m <= 20000 else 3):
                t0 = time.perf_counter(); ctx.psrf_eval(burn, start, m, 1, meth); t.append(time.perf_counter() - t0)
            row[f"gpu_psrf_{name}_ms"] = round(1e3 * float(np.median(t)), 3)
        ctx.ess_eval(burn, m, [1, 2, 3])
        t = []
        for _ in range(10):
            t0 = time.perf_counter(); ctx.ess_eval(burn, m, [1, 2, 3]); t.append(time.perf_counter() - t0)
        row["gpu_ess_3cols_ms"] = round(1e3 * float(np.median(t)), 3)
    if m <= 5000:  # CPU oracle, one uncached check (the reference caches distances between checks: its steady-state
        t0 = time.perf_counter()  # cost per check is ~4m new distances + 4m^2 cache reads)
        ts.psrf_rows_mean(0, 1, start, burn, m, 1); ts.psrf_rows_mean(1, 0, start, burn, m, 1)
        row["cpu_psrf_uncached_ms"] = round(1e3 * (time.perf_counter() - t0), 3)
    t0 = time.perf_counter()
    O.trace_ess_converged([x[0], x[1]], burn, m, [1, 2, 3])
    row["cpu_ess_3cols_ms"] = round(1e3 * (time.perf_counter() - t0), 3)
    res.append(row)
print(json.dumps(res, indent=1))
