#!/usr/bin/env python
"""bench.py -- throughput of the ASM convergence-diagnostic hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload psrf|ess|cfg5] [--method fp4|i8|popc]

Prints ONE JSON line (rank 0).  A "step" is one full pass of the hot path over one batch of synthetic input:

  psrf (default; BASELINE.json configs[1], "cfg2"): 4 chains x 10,000 trees x 100 taxa, all-pairs RF + the (d+1)^2
      within/between-chain sums + per-row PSRF + per-pair means, i.e. everything TreePSRF.converged2sided
      (TreePSRF.java:141-158) asks of the device for every ordered chain pair.  metric = tree-pairs/sec, counting the
      ordered (row tree, column tree) pairs whose distance enters the sums (T*T per step; the GPU evaluates each
      unordered pair once).  At N > 1 GPUs per-GPU work is held fixed (trees per chain scale with sqrt(N)):
      "scaling": "weak".
  The same line carries, at every N, two sub-objects for the other shapes the north star names:
      "ess"  (configs[2], "cfg3"): 1,000 traces x 1,000,000 samples per GPU, TraceESS.ACT/calcESS per trace (weak);
      "cfg5" (configs[4]): 32 chains x 50,000 trees x 500 taxa, the SAME input at every N (strong scaling), with an
             `S_checksum` of the full int64 table that must be equal across N.

How the N GPUs are driven (both inside libasmgpu.so, asm_b200/csrc/comm.cu; no data-path collective in Python):
  * under torchrun (one process per GPU): every rank creates a one-GPU context and joins an in-library NCCL
    communicator (asm_comm_init_rank; torch.distributed/gloo only carries the 128-byte id, barriers and the
    max-over-ranks of the timings);
  * plain `python bench.py --gpus N` (no torchrun): ONE process drives the N GPUs through asm_create_multi, the way
    the reference's single caller thread would.
The timer is the same at every N: CUDA events on the library's stream(s), max over GPUs.

`value` is measured with the inputs resident in HBM; `e2e` goes through the C ABI with host buffers (pinned),
host->device and device->host copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(n_chains=4, n_trees=10000, n_taxa=100, beta=3.3, moves=2, start_seed=1, seed0=20240001)
CFG5 = dict(n_chains=32, n_trees=50000, n_taxa=500, beta=3.9, moves=2, start_seed=1, seed0=20240100)
CFG3 = dict(n_traces=1000, n_samples=1_000_000, seed0=777)
PHIS, MUS = [0.0, 0.5, 0.9, 0.99], [0.0, 1.0, -1e4]


def measured_traffic(key):
    """DRAM bytes per launch of the named kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def live_peaks(device_index):
    """Denominators measured in THIS run (tools/peaks.py): vendor-library GEMM throughput for the tensor-core kernels,
    pipe issue rates for the CUDA-core kernels."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import peaks as P
    out = {}
    try:
        out["gemm"] = P.gemm_peaks(f"cuda:{device_index}", sustain_s=1.0)
    except Exception as e:  # noqa: BLE001
        out["gemm"] = {"error": repr(e)[:300]}
    out["pipes"] = P.pipe_peaks(device_index)
    return out


def tensor_peak(live, peaks, kind, sustained):
    """(peak TFLOP/s, description) for the roofline of a tensor-core kernel: the vendor library's GEMM of the same
    operand type measured in this run; MEASURED_PEAKS.json's bf16 times the datapath ratio only if that failed."""
    which = "sustained" if sustained else "burst"
    g = (live or {}).get("gemm", {})
    keys = {"fp4": ("fp4_mx_e8m0_block32", "fp4_nv_e4m3_block16"), "i8": ("int8",)}[kind]
    cands = [(g[k][which], k) for k in keys if isinstance(g.get(k), dict)]
    if cands:
        v, k = max(cands)
        n = g[k].get("n", g.get("n", 8192))
        return v, f"cuBLASLt {k} GEMM through torch (best of the sizes tried; burst at {n}^3), {which}, measured in this run (tools/peaks.py)"
    mult = 4.0 if kind == "fp4" else 2.0
    base = peaks["bf16_tflops_sustained"] if sustained else peaks["bf16_tflops"]
    return mult * base, (f"{mult:g} x bf16 {which} of {peaks['source']} (DERIVED: the {kind} GEMM could not be measured here: "
                         f"{[g.get(k + '_error') for k in keys]})")


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
def gen_psrf_input(cfg, n_trees, want_topologies=False):
    """Synthetic chains -> (encoder, [bits per chain], [topologies per chain] or None).

    Chains longer than cfg["n_trees"] (the weak-scaling runs at N > 1 GPUs) keep the first cfg["n_trees"]
    generated trees and append a seeded resample of them, so the clade dictionary -- the K dimension of the
    contraction, i.e. the work per tree pair -- is the same at every N."""
    import synthgen as G
    enc = G.Encoder(cfg["n_taxa"])  # the generator's own dictionary: input generation uses nothing of libasmgpu.so
    n_base = min(n_trees, cfg["n_trees"])
    ids, topo = [], []
    for c in range(cfg["n_chains"]):
        l, r, root, i = G.synth_tree_chain(cfg["n_taxa"], n_base, cfg["start_seed"], cfg["seed0"] + c, cfg["moves"],
                                           cfg["beta"], topologies=want_topologies, encoder=enc)
        if n_trees > n_base:
            pick = np.random.default_rng(cfg["seed0"] + 1000 + c).integers(0, n_base, n_trees - n_base)
            i = np.concatenate([i, i[pick]])
            if want_topologies:
                l, r, root = np.concatenate([l, l[pick]]), np.concatenate([r, r[pick]]), np.concatenate([root, root[pick]])
        ids.append(i)
        topo.append((l, r, root))
    words = enc.words
    bits = [enc.bits(i, words) for i in ids]
    enc.chain_ids = ids  # the clade ids the rows were built from (what the end-to-end leg uploads)
    return enc, bits, (topo if want_topologies else None)


def gen_traces(n_traces, n_samples, seed0, first=0, out=None):
    from concurrent.futures import ThreadPoolExecutor
    import synthgen as G
    x = np.empty((n_traces, n_samples), dtype=np.float64) if out is None else out

    def one(j):
        g = first + j
        G.synth_ar1(n_samples, MUS[g % 3], PHIS[g % 4], seed0 + g, out=x[j])

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(n_traces)))
    return x


def pinned_like(a: np.ndarray):
    """Copy `a` into page-locked host memory (torch is plumbing here) and return (tensor, ndarray view)."""
    import torch
    t = torch.from_numpy(a)
    try:
        t = t.pin_memory()
    except Exception:
        pass
    return t, t.numpy()


# --------------------------------------------------------------------------------------------------
def oracle_tree_set(cfg, n_trees):
    import oracle as O
    import synthgen as G
    ts = O.TreeSet(cfg["n_chains"], cfg["n_taxa"])
    for c in range(cfg["n_chains"]):
        l, r, root, _ = G.synth_tree_chain(cfg["n_taxa"], n_trees, cfg["start_seed"], cfg["seed0"] + c, cfg["moves"],
                                           cfg["beta"])
        ts.add(c, l, r, root)
    return ts


def cpu_baseline_psrf(cfg, n_trees, sample_rows=2048, ts=None):
    """The reference's CPU path for this workload, as restated by the oracle (Java cannot run here):
    calcPSRF's two column loops (TreePSRF.java:274-286) for `sample_rows` row trees against every
    column tree of every chain, OpenMP over rows."""
    import oracle as O
    cores = O.use_all_cores()
    if ts is None:
        ts = oracle_tree_set(cfg, n_trees)
    rows = min(sample_rows, n_trees)
    burnin = [0] * cfg["n_chains"]
    T = cfg["n_chains"] * n_trees
    t0 = time.perf_counter()
    # rows [n_trees-rows, n_trees) of every chain against all columns of all chains
    S = ts.psrf_sums(burnin, n_trees - rows, n_trees, 1)
    dt_all = time.perf_counter() - t0
    pairs = cfg["n_chains"] * rows * T
    return {"value": pairs / dt_all, "unit": "tree-pairs/s", "cores": cores, "kind": "port",
            "sample": f"{cfg['n_chains']}x{rows} row trees x {T} column trees ({pairs:.3g} ordered pairs) in {dt_all:.2f}s; "
                      f"C restatement of TreePSRF.calcPSRF with set-intersection RF, OpenMP", "seconds": dt_all,
            "checksum": float(S.sum())}


def cpu_baseline_ess(n_samples, seed0, budget_traces=None):
    import oracle as O
    cores = O.use_all_cores()
    n_tr = budget_traces or max(cores, 4)
    x = gen_traces(n_tr, n_samples, seed0)
    t0 = time.perf_counter()
    act, ess = O.ess_batch(x, literal=False)
    dt = time.perf_counter() - t0
    out = {"value": n_tr * n_samples / dt, "unit": "trace-samples/s", "cores": cores, "kind": "port",
           "sample": f"{n_tr} traces x {n_samples} samples in {dt:.2f}s; C restatement of TraceESS.ACT, closed-form final "
                     f"pass (the literal per-sample recomputation the JVM executes is ~6x slower), OpenMP over traces",
           "seconds": dt}
    return out


# --------------------------------------------------------------------------------------------------
class Job:
    """How this process takes part in the N-GPU job (see the module docstring)."""

    def __init__(self, n_gpus: int):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.group_devices = None
        self.numa_cpus = 0
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("gloo")  # ids, barriers, maxima of timings; the data path is NCCL inside the library
            self.dist = dist
            self.n_gpus = self.world
            self.launch = "torchrun: one process per GPU, in-library NCCL communicator (asm_comm_init_rank)"
            # host buffers this rank allocates and pins from here on sit on its GPU's NUMA node (a no-op where the
            # platform shows a single node)
            from asm_b200 import _lib as L
            self.numa_cpus = int(L.lib().asm_bind_thread_near_device(self.local))
        elif n_gpus > 1:
            self.group_devices = list(range(n_gpus))
            self.n_gpus = n_gpus
            self.launch = "one process drives all GPUs (asm_create_multi: ncclCommInitAll + one host thread per GPU)"
        else:
            self.n_gpus = 1
            self.launch = "one process, one GPU"

    def make_ctx(self, n_chains):
        import asm_b200 as A
        if self.group_devices is not None:
            return A.GpuContext(n_chains, devices=self.group_devices)
        ctx = A.GpuContext(n_chains, device=self.local)
        if self.world > 1:
            from asm_b200 import multi
            multi.init_comm(ctx, self.dist)
        return ctx

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.sync()
        if self.dist is not None:
            self.dist.barrier()

    def reduce(self, x: float, op: str = "max") -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def timed_steps(job, ctx, step, steps, flush=True):
    """`steps` device-timed steps: CUDA events on the library's stream(s) around each step (max over the GPUs of a
    group inside the library), barrier + sync before each, mean over steps, then the max over ranks."""
    per = []
    out = None
    for _ in range(steps):
        if flush:
            ctx.flush_l2()  # outside the timed bracket
        job.barrier(ctx)
        ctx.timer_mark(0)
        out = step()
        ctx.timer_mark(1)
        per.append(max(ctx.timer_elapsed_ms(0, 1), 0.0))
    return job.reduce(float(np.mean(per)), "max"), out


def s_checksum(S: np.ndarray) -> str:
    """Sum of the int64 table modulo 2^64: equal across GPU counts iff (up to collisions) the tables are."""
    return format(int(np.ascontiguousarray(S).view(np.uint64).sum(dtype=np.uint64)), "016x")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Java itself
    cannot run in this image), bounded sample per step, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    import oracle as O
    O.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm uses every host core it may
    if args.workload == "ess":
        n = CFG3["n_samples"]
        vals = []
        for s in range(warm + steps):
            r = cpu_baseline_ess(n, CFG3["seed0"])
            if s >= warm:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        line = {"metric": "trace-samples/sec (ESS)", "value": v, "unit": "trace-samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "impl": "reference", "config": {"workload": "cfg3: 1000 traces x 1M samples (bounded sample per step)"},
                "cpu_baseline": {k: vals[-1][k] for k in ("kind", "cores", "sample")} | {"value": v, "unit": "trace-samples/s"},
                "e2e": {"value": v, "unit": "trace-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        cfg = CFG5 if args.workload == "cfg5" else CFG2
        n_trees = cfg["n_trees"] if args.workload == "cfg5" else int(round(cfg["n_trees"] * math.sqrt(args.gpus)))
        vals = []
        ts = oracle_tree_set(cfg, n_trees)  # built once; a step is the calcPSRF loops over a bounded block of rows
        for s in range(warm + steps):
            r = cpu_baseline_psrf(cfg, n_trees, sample_rows=args.ref_rows, ts=ts)
            if s >= warm:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        line = {"metric": "tree-pairs/sec (RF+PSRF)", "value": v, "unit": "tree-pairs/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/int64", "data": "synthetic",
                "impl": "reference",
                "config": {"workload": f"{'cfg5' if args.workload == 'cfg5' else 'cfg2'}: {cfg['n_chains']} chains x {n_trees} trees x "
                                       f"{cfg['n_taxa']} taxa (bounded row sample per step)"},
                "cpu_baseline": {k: vals[-1][k] for k in ("kind", "cores", "sample")} | {"value": v, "unit": "tree-pairs/s"},
                "e2e": {"value": v, "unit": "tree-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    with open("/proc/self/maps") as f:  # evidence: this arm maps the oracle and the input generators, not the product
        line["native_libs"] = sorted({os.path.relpath(l.split()[-1], ROOT) for l in f if l.rstrip().endswith(".so") and ROOT in l})
    _emit(line)


# --------------------------------------------------------------------------------------------------
def bench_psrf(job, args, which, method_name, steps, warmup, live, others=False):
    """One PSRF workload (`which`: "cfg2" weak-scaled, or "cfg5" strong) -> the dict of its bench line."""
    from asm_b200 import _lib as L
    peaks = measured_peaks()
    world = job.n_gpus
    cfg = CFG5 if which == "cfg5" else CFG2
    n_trees = cfg["n_trees"] if which == "cfg5" else int(round(cfg["n_trees"] * math.sqrt(world)))
    method = {"i8": L.ASM_RF_I8, "fp4": L.ASM_RF_FP4, "popc": L.ASM_RF_POPC}[method_name]
    C = cfg["n_chains"]
    T = C * n_trees
    enc, bits, _ = gen_psrf_input(cfg, n_trees)
    D = enc.num_clades
    burnin = [0] * C
    ctx = job.make_ctx(C)
    pinned = [pinned_like(b) for b in bits]
    del bits
    # end-to-end leg: the host hands over what its encoder emits, uint16 clade ids (198 bytes per 100-taxon tree against
    # 352 bytes of bit row); the library builds the bit rows on the GPU (asm_trees_set_ids)
    # (at N > 1 the sharded upload of finished rows -- each GPU 1/N of them over PCIe, one all-gather over NVLink -- moves
    # fewer bytes per GPU than every GPU taking all the ids, so the multi-GPU lines keep it)
    use_ids = D <= 0xFFFF and not args.e2e_bits and world == 1
    pinned_ids = [pinned_like(i.astype(np.uint16)) for i in enc.chain_ids] if use_ids else None
    h2d_bytes = sum(p[1].nbytes for p in (pinned_ids if use_ids else pinned))

    def upload_bits():  # at N > 1 every row crosses PCIe once per job: GPU r uploads the r-th slice, one all-gather replicates it
        for c in range(C):
            ctx.trees_set(c, pinned[c][1])

    def upload():
        if not use_ids:
            return upload_bits()
        for c in range(C):  # at N > 1 every GPU takes the ids itself
            ctx.trees_set_ids(c, pinned_ids[c][1], D)

    def step():  # sums on every GPU's share of the tile pairs, ONE int64 all-reduce inside the library, finish
        return ctx.psrf_eval(burnin, 0, n_trees, 1, method)

    upload_bits()
    for _ in range(warmup):
        mean = step()
        ctx.flush_l2()
    # ---- timed: inputs resident in HBM --------------------------------------------------------------
    ctx.profile_enable(False)
    sampler = ClockSampler(job.local, period_s=0.001 if which == "cfg2" else 0.02)
    job.barrier(ctx)
    launches0 = ctx.launch_count()
    sampler.start()
    ms, mean = timed_steps(job, ctx, step, steps)
    clocks = sampler.stop()
    n_local = max(1, len(ctx.devices))
    launches = job.reduce((ctx.launch_count() - launches0 - steps * n_local) / max(steps, 1), "sum")  # minus the flush kernels
    value = T * T / (ms * 1e-3)

    # ---- kernel-only time for the roofline (CUDA events around every launch, same stream) --------------
    ctx.profile_enable(True)
    ctx.profile_reset()
    n_prof = 1 if which == "cfg5" else max(3, min(steps, 10))
    for _ in range(n_prof):
        ctx.flush_l2()
        step()
    fam = L.K_RF_POPC if method == L.ASM_RF_POPC else L.K_RF_I8  # both tensor-core kernels report under K_RF_I8
    # profile_get: time of the slowest GPU of a group; one distance-kernel launch per GPU and step
    fam_ms = {name: job.reduce(ctx.profile_get(f)[0] / n_prof, "max") for name, f in
              [("gather", L.K_GATHER), ("rf", fam), ("finish", L.K_PSRF_FINISH)]}
    ctx.profile_enable(False)
    k_ms = fam_ms["rf"]
    lay = ctx.last_layout()
    Dp, Tp = lay["padded_bits"], lay["padded_rows"]
    sustained = which == "cfg5"  # a seconds-long kernel runs at the sustained clock, a 0.6 ms one at the burst clock
    if method in (L.ASM_RF_I8, L.ASM_RF_FP4):
        kind = "fp4" if method == L.ASM_RF_FP4 else "i8"
        alg = D * T * (T - 1) / world  # per GPU: its share of the unordered pairs, 2 flop per MAC
        achieved = alg / (k_ms * 1e-3) / 1e12
        peak, src = tensor_peak(live, peaks, kind, sustained)
        nominal = 9000.0 if kind == "fp4" else 4500.0
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": measured_traffic(("rf_f4x2_kernel" if kind == "fp4" else "rf_i8x2_kernel") + "@cfg2")
                if (world == 1 and which == "cfg2") else None,
                "kernel": "rf_f4x2_kernel" if kind == "fp4" else "rf_i8x2_kernel", "kernel_ms": k_ms, "peak_source": src,
                "frac_of_nominal": achieved / nominal, "nominal_tflops": nominal,
                "executed_over_algorithmic": (Tp * (Tp + 256) / 2.0 * 2 * Dp) / (D * T * (T - 1.0)),
                "algorithmic": "D*T*(T-1)/N flop per launch and GPU (unordered pairs, 2 flop per MAC)" +
                               ("; e2m1 indicators, unit UE8M0 scales, f32 accumulators: exact" if kind == "fp4" else "")}
    else:
        pipes = (live or {}).get("pipes", {})
        words_pairs = (D / 64.0) * T * (T - 1) / 2 / world
        achieved = words_pairs * 2 / (k_ms * 1e-3)  # POPC lane-ops/s: one 64-bit popcount = 2 POPC
        peak = pipes.get("popc", {}).get("ops_per_s")
        roof = {"bound": "int", "achieved": achieved / 1e12, "peak": (peak or 4.45e12) / 1e12, "unit": "Tpopc32/s",
                "frac": achieved / (peak or 4.45e12), "traffic": measured_traffic("rf_popc_psrf_kernel@cfg2") if world == 1 else None,
                "kernel": "rf_popc_psrf_kernel", "kernel_ms": k_ms,
                "peak_source": "POPC issue rate, tools/microbench measured in this run" if peak else "POPC issue rate of profiles/r01_microbench_pipes.json",
                "algorithmic": "(D/64)*T*(T-1)/2 64-bit XOR-popcounts (2 POPC each) per launch; integer-pipe bound, not HBM",
                "hbm_gbs": (Tp * Dp / 8 + 8 * Tp * C) / world / (k_ms * 1e-3) / 1e9}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ------------------------
    e2e_steps = []
    e2e_warm = 0 if which == "cfg5" else warmup
    for i in range(e2e_warm + (1 if which == "cfg5" else steps)):
        ctx.flush_l2()
        job.barrier(ctx)
        t0 = time.perf_counter()
        upload()
        mean_e = step()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= e2e_warm:
            e2e_steps.append(dt)
    e_ms = job.reduce(float(np.mean(e2e_steps)), "max")
    e2e = {"value": T * T / (e_ms * 1e-3), "unit": "tree-pairs/s", "h2d_bytes_per_step": int(h2d_bytes),
           "d2h_bytes_per_step": int(C * C * 8), "ms_per_step": e_ms,
           "timer": "host clock around asm_trees_set[_ids] x C + asm_psrf_eval, max over ranks",
           "bit_identical_to_resident": bool(np.array_equal(np.asarray(mean_e), np.asarray(mean), equal_nan=True))}
    S = ctx.psrf_sums(burnin, 0, n_trees, 1, method)  # the complete int64 table (all-reduced at N > 1)
    checksum = s_checksum(S)
    del S
    line = {"metric": "tree-pairs/sec (RF+PSRF)", "value": value, "unit": "tree-pairs/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if which == "cfg5" else "weak",
            "vs_baseline": None, "dtype": {L.ASM_RF_I8: "int8->s32 (tcgen05), int64 sums, f64 psrf", L.ASM_RF_FP4: "e2m1->f32 (tcgen05 mxf4, exact on 0/1), int64 sums, f64 psrf", L.ASM_RF_POPC: "u64 popc, int64 sums, f64 psrf"}[method],
            "data": "synthetic",
            "config": {"workload": f"{which}: {C} chains x {n_trees} trees x {cfg['n_taxa']} taxa, burnin=0 smoothing=1 delta=1",
                       "method": method_name, "clades_D": int(D), "padded_bits": int(Dp), "padded_rows": int(Tp), "trees_T": int(T),
                       "launch": job.launch,
                       "l2": "L2 flushed (256 MiB write) between timed steps, outside the timed bracket",
                       "timing": "CUDA events on the library stream(s) per step, max over GPUs (the same timer at every N)",
                       "generator": {k: cfg[k] for k in ("beta", "moves", "start_seed", "seed0")},
                       "weak_scaling": ("trees per chain = 10000*sqrt(N); trees beyond the first 10000 are a seeded resample of them, so D "
                                        "(work per pair) is the same at every N") if which == "cfg2" else None,
                       "e2e_upload": ("asm_trees_set_ids from pinned uint16 clade ids (the encoder's output), bit rows built on the GPU" +
                                      ("" if world == 1 else "; every GPU takes the ids itself (h2d_bytes_per_step is one GPU's)"))
                       if use_ids else
                                     ("asm_trees_set from pinned host rows" + ("" if world == 1 else
                                      ": GPU r uploads the r-th slice of every chain, one in-library ncclAllGather replicates it "
                                      "(h2d_bytes_per_step is the whole job's)")),
                       "psrf_mean_01": float(mean[0, 1]), "psrf_mean_10": float(mean[1, 0])},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(round(launches)), "roofline": roof,
            "step_breakdown_ms": fam_ms, "S_checksum": checksum}
    # ---- the other ways the north star names (popcount-XOR, int8 tcgen05), same inputs, same step ---------------------
    if others:
        res = {}
        for name, m in (("i8", L.ASM_RF_I8), ("fp4", L.ASM_RF_FP4), ("popc", L.ASM_RF_POPC)):
            if m == method:
                continue
            n_steps = 3 if m == L.ASM_RF_POPC else 10
            for _ in range(3):
                mean_m = ctx.psrf_eval(burnin, 0, n_trees, 1, m)
            ts = []
            for _ in range(n_steps):
                ctx.flush_l2()
                ctx.timer_mark(0)
                mean_m = ctx.psrf_eval(burnin, 0, n_trees, 1, m)
                ctx.timer_mark(1)
                ts.append(ctx.timer_elapsed_ms(0, 1))
            ctx.profile_enable(True)
            ctx.profile_reset()
            for _ in range(3):
                ctx.flush_l2()
                ctx.psrf_eval(burnin, 0, n_trees, 1, m)
            km, kn = ctx.profile_get(L.K_RF_POPC if m == L.ASM_RF_POPC else L.K_RF_I8)
            ctx.profile_enable(False)
            km /= max(kn, 1)
            ms_m = float(np.mean(ts))
            entry = {"ms_per_step": ms_m, "value": T * T / (ms_m * 1e-3), "unit": "tree-pairs/s", "kernel_ms": km,
                     "bit_identical_means": bool(np.array_equal(np.asarray(mean_m), np.asarray(mean), equal_nan=True))}
            if m == L.ASM_RF_POPC:
                pk = (live or {}).get("pipes", {}).get("popc", {}).get("ops_per_s")
                entry["xor_popc64_per_s"] = (D / 64.0) * T * (T - 1) / 2 / (km * 1e-3)
                entry["frac_of_popc_issue_rate"] = 2 * entry["xor_popc64_per_s"] / pk if pk else None
            else:
                pk, src = tensor_peak(live, peaks, name, False)
                entry["algorithmic_tflops"] = D * T * (T - 1) / (km * 1e-3) / 1e12
                entry["frac"] = entry["algorithmic_tflops"] / pk
                entry["peak"], entry["peak_source"] = pk, src
            res[name] = entry
        line["other_methods"] = res
    ctx.close()
    return line


def bench_ess(job, args, steps, warmup, live):
    from asm_b200 import _lib as L
    import torch
    world = job.n_gpus
    n_tr = args.ess_traces or CFG3["n_traces"]  # per GPU
    n = CFG3["n_samples"]
    # traces this PROCESS holds: its own GPU's (torchrun) or every GPU's (one process drives them all)
    n_here = n_tr * (len(job.group_devices) if job.group_devices else 1)
    first = job.rank * n_tr + args.ess_first
    ctx = job.make_ctx(1)
    host_t = torch.empty((n_here, n), dtype=torch.float64)
    try:
        host_t = host_t.pin_memory()
    except Exception:
        pass
    host = host_t.numpy()
    gen_traces(n_here, n, CFG3["seed0"], first=first, out=host)
    ctx.ess_traces_put(host)  # resident: a group deals the traces to its GPUs in contiguous blocks

    def step():  # per-GPU evaluation + the ESS vector of the whole job on every rank (in-library all-gather)
        act, ess = ctx.ess_traces_eval()
        return act, (ctx.all_gather_f64(ess) if job.dist is not None else ess)

    for _ in range(warmup):
        step()
    sampler = ClockSampler(job.local, period_s=0.05)
    job.barrier(ctx)
    l0 = ctx.launch_count()
    sampler.start()
    ms, (act, ess_all) = timed_steps(job, ctx, step, steps, flush=False)  # 8 GB of traces per GPU and step: larger than L2
    clocks = sampler.stop()
    launches = job.reduce((ctx.launch_count() - l0) / max(steps, 1), "sum")
    value = world * n_tr * n / (ms * 1e-3)

    def profiled():
        ctx.profile_enable(True)
        ctx.profile_reset()
        ctx.ess_traces_eval()
        k, kn = ctx.profile_get(L.K_ESS_LAG)
        f, _ = ctx.profile_get(L.K_ESS_FINISH)
        ctx.profile_enable(False)
        return job.reduce(k, "max"), job.reduce(f, "max"), job.reduce(ctx.ess_last_lag_count(), "sum")

    k_ms, f_ms, slots = profiled()          # as shipped: staged lag evaluation
    ctx.ess_set_adaptive(False)
    job.barrier(ctx)
    ctx.timer_mark(2)
    act_full, ess_full = ctx.ess_traces_eval()
    ctx.timer_mark(3)
    full_step_ms = job.reduce(ctx.timer_elapsed_ms(2, 3), "max")
    kf_ms, ff_ms, slots_full = profiled()   # every lag of every trace (what the reference evaluates)
    ctx.ess_set_adaptive(True)
    lmax = min(n, 2000)
    flops_full = 2.0 * (lmax * n - lmax * (lmax - 1) / 2) * n_tr * world
    flops = 2.0 * slots * n  # (trace, lag) chains evaluated x n mul+add pairs (upper bound: ignores the i < lag head)
    pipes = (live or {}).get("pipes", {})
    fp64_peak = pipes.get("fp64_issue_tflops")
    peak_src = "DMUL+DADD issue rate, tools/microbench.cu measured in this run" if fp64_peak else \
        "DMUL+DADD issue rate of profiles/r01_microbench_pipes.json (tools/microbench not runnable here)"
    fp64_peak = fp64_peak or 18.38
    ach = flops / world / (k_ms * 1e-3) / 1e12
    ach_full = flops_full / world / (kf_ms * 1e-3) / 1e12
    mine = ess_all[first:first + n_here] if job.dist is not None else ess_all
    roof = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
            "traffic": measured_traffic("ess_lag_kernel@cfg3_staged_5_launches") if world == 1 else None,
            "traffic_note": "sum over the staged launches of one step (each stage re-reads its open traces); the HBM floor of 8 B/sample is not binding",
            "kernel": "ess_lag_kernel", "kernel_ms": k_ms, "finish_ms": f_ms, "peak_source": peak_src,
            "algorithmic": "2*n flop per evaluated (trace, lag) chain, i.e. EXECUTED work: separate DMUL and DADD (bit-exact, no FMA), so "
                           "the ceiling is the FP64 issue rate (half the DFMA rating); the staged evaluation stops a trace once the "
                           "integral loop of TraceESS.java:161-172 has stopped",
            "lag_chains_evaluated": int(slots), "lag_chains_all": int(n_tr) * 2048 * world,
            "hbm_gbs": 8.0 * n_tr * n / (k_ms * 1e-3) / 1e9,
            "all_lags": {"step_ms": full_step_ms, "kernel_ms": kf_ms, "achieved": ach_full, "frac": ach_full / fp64_peak,
                         "value": world * n_tr * n / (full_step_ms * 1e-3),
                         "bit_identical_to_staged": bool(np.array_equal(ess_full, mine, equal_nan=True)),
                         "algorithmic": "2*(L*n - L(L-1)/2) flop per trace, L = 2000: every lag, what the reference evaluates"}}
    e2e_ms = []
    for i in range(2):
        job.barrier(ctx)
        t0 = time.perf_counter()
        act2, ess2 = ctx.ess_batch(host)
        if job.dist is not None:
            ess2 = ctx.all_gather_f64(ess2)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    e_ms = job.reduce(float(min(e2e_ms)), "max")
    # ---- the same 1,000 traces as the one-GPU cfg3 run, dealt over the N GPUs (BASELINE.json configs[2]: "1 and 8 GPUs") ----
    strong = None
    if world > 1 and n_tr % world == 0:
        per = n_tr // world
        if job.dist is not None:   # torchrun: this rank holds block `rank` of traces 0 .. n_tr-1
            gen_traces(per, n, CFG3["seed0"], first=job.rank * per, out=host[:per])
            ctx.ess_traces_put(host[:per])
        else:                      # one process: the group deals the blocks itself
            gen_traces(n_tr, n, CFG3["seed0"], first=0, out=host[:n_tr])
            ctx.ess_traces_put(host[:n_tr])
        for _ in range(warmup):
            step()
        job.barrier(ctx)
        s_ms, (_, s_ess) = timed_steps(job, ctx, step, steps, flush=False)
        strong = {"scaling": "strong", "traces_total": n_tr, "traces_per_gpu": per, "ms_per_step": s_ms,
                  "value": n_tr * n / (s_ms * 1e-3), "unit": "trace-samples/s",
                  "ess_checksum": format(int(np.ascontiguousarray(s_ess).view(np.uint64).sum(dtype=np.uint64)), "016x"),
                  "note": f"the {n_tr} traces of the one-GPU cfg3 line (same ess_checksum) in contiguous blocks of {per} per GPU; "
                          "the fewer traces a GPU holds, the more of its rounds are one latency-bound walk of a chain over all samples"}
    from asm_b200.multi import java_min
    line = {"metric": "trace-samples/sec (ESS)", "value": value, "unit": "trace-samples/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg3: {n_tr} traces x {n} samples per GPU, AR(1) phi in {PHIS}, mu in {MUS}, max_lag 2000",
                       "l2": "inputs (8 GB per GPU and step) larger than L2", "launch": job.launch,
                       "host_binding": (f"rank bound to the {job.numa_cpus} CPUs local to its GPU before allocating" if job.numa_cpus
                                        else "none (one NUMA node, or the platform hides the topology)"),
                       "timing": "CUDA events on the library stream(s) per step, max over GPUs"},
            "clocks": clocks, "gpu_launches": int(round(launches)), "roofline": roof,
            "e2e": {"value": world * n_tr * n / (e_ms * 1e-3), "unit": "trace-samples/s",
                    "h2d_bytes_per_step": int(host.nbytes) * (job.world if job.dist is not None else 1),
                    "d2h_bytes_per_step": int(16 * n_tr * world), "ms_per_step": e_ms},
            "min_ess": java_min(ess_all),
            "ess_checksum": format(int(np.ascontiguousarray(ess_all).view(np.uint64).sum(dtype=np.uint64)), "016x"),
            "bit_identical_e2e_vs_resident": bool(np.array_equal(ess_all, ess2, equal_nan=True))}
    if strong is not None:
        line["strong"] = strong
    ctx.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="psrf", choices=["psrf", "ess", "cfg5"])
    ap.add_argument("--method", default=os.environ.get("ASM_BENCH_METHOD", "fp4"), choices=["i8", "fp4", "popc"])
    ap.add_argument("--cpu-rows", type=int, default=2048, help="row trees per chain in the CPU-baseline sample")
    ap.add_argument("--ref-rows", type=int, default=512,
                    help="row trees per chain per step of --impl reference (about 2 s of CPU work per step at cfg2)")
    ap.add_argument("--ess-traces", type=int, default=0, help="traces per GPU of the ESS block (default 1000 = cfg3)")
    ap.add_argument("--ess-first", type=int, default=0, help="index of the first synthetic trace (probe: the traces another rank would hold)")
    ap.add_argument("--e2e-bits", action="store_true", help="end-to-end leg uploads finished bit rows (asm_trees_set) instead of clade ids")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true", help="skip the ESS (cfg3) block")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the cfg5 strong-scaling block")
    ap.add_argument("--no-other-methods", action="store_true", help="skip the int8 / popcount comparison runs")
    ap.add_argument("--no-live-peaks", action="store_true", help="skip the in-run GEMM / pipe peak measurements")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version there when
    # NCCL_DEBUG asks for it), so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    globals()["_emit"] = emit
    if args.impl == "reference":
        run_reference(args)
        return
    job = Job(args.gpus)
    t_start = time.time()
    # roofline denominators measured in this lease, on this process's GPU, before the workload (every rank measures its
    # own GPU so that all GPUs are in the same thermal state; rank 0's numbers are reported)
    live = None if args.no_live_peaks else live_peaks(job.local)
    job.barrier()
    if args.workload == "ess":
        line = bench_ess(job, args, args.steps, args.warmup, live)
        if job.rank == 0 and job.n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_ess(CFG3["n_samples"], CFG3["seed0"])
    elif args.workload == "cfg5":
        line = bench_psrf(job, args, "cfg5", args.method, min(args.steps, 2), 1, live)
    else:
        line = bench_psrf(job, args, "cfg2", args.method, args.steps, args.warmup, live,
                          others=job.n_gpus == 1 and not args.no_other_methods)
        if job.rank == 0 and job.n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_psrf(CFG2, CFG2["n_trees"], sample_rows=args.cpu_rows)
        if not args.no_ess:
            ess_line = bench_ess(job, args, min(args.steps, 5), 3, live)
            if job.rank == 0 and job.n_gpus == 1 and not args.no_cpu_baseline:
                ess_line["cpu_baseline"] = cpu_baseline_ess(CFG3["n_samples"], CFG3["seed0"])
            line["ess"] = ess_line
        if not args.no_cfg5:
            # strong scaling of the north star's largest shape: the same 1.6M trees at every N; a step is seconds long,
            # so one warm-up and at most two timed steps
            line["cfg5"] = bench_psrf(job, args, "cfg5", args.method, min(args.steps, 2), 1, live)
    if live is not None:
        line["live_peaks"] = live
    line["wall_s"] = round(time.time() - t_start, 1)
    if job.rank == 0:
        _emit(line)
    job.close()


if __name__ == "__main__":
    main()
