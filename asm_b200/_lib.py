"""ctypes binding of include/asmgpu.h.  Plumbing only: every call goes straight to libasmgpu.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasmgpu.so")

ASM_OK, ASM_E_INVALID, ASM_E_NO_DEVICE, ASM_E_CUDA, ASM_E_RANGE, ASM_E_NOMEM, ASM_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
ASM_RF_POPC, ASM_RF_I8, ASM_RF_FP4 = 0, 1, 2
ASM_MAX_LAG = 2000
K_GATHER, K_RF_POPC, K_RF_I8, K_PSRF_FINISH, K_CLADE_COUNTS, K_ESS_LAG, K_ESS_FINISH, K_MISC = range(8)
_ERR_NAMES = {-1: "ASM_E_INVALID", -2: "ASM_E_NO_DEVICE", -3: "ASM_E_CUDA", -4: "ASM_E_RANGE", -5: "ASM_E_NOMEM",
              -6: "ASM_E_UNSUPPORTED"}


class AsmGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Shard(C.Structure):
    _fields_ = [("shard_index", C.c_int32), ("shard_count", C.c_int32)]


_p = C.c_void_p
_i32, _i64, _u64, _f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double
_i32p, _i64p, _f64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> (restype, argtypes); exactly the declarations of include/asmgpu.h
SIGNATURES = {
    "asm_version": (C.c_char_p, []),
    "asm_device_count": (_i32, []),
    "asm_create": (_i32, [C.POINTER(_p), _i32, _i32]),
    "asm_destroy": (_i32, [_p]),
    "asm_create_multi": (_i32, [C.POINTER(_p), _p, _i32, _i32]),
    "asm_comm_unique_id": (_i32, [_p]),
    "asm_comm_init_rank": (_i32, [_p, _p, _i32, _i32]),
    "asm_comm_info": (_i32, [_p, _i32p, _i32p, _i32p]),
    "asm_comm_nccl_version": (_i32, []),
    "asm_comm_all_gather_f64": (_i32, [_p, _p, _i64, _p]),
    "asm_last_error": (C.c_char_p, [_p]),
    "asm_trees_set": (_i32, [_p, _i32, _i64, _i32, _p]),
    "asm_trees_append": (_i32, [_p, _i32, _i64, _i32, _p]),
    "asm_trees_count": (_i32, [_p, _i32, _i64p, _i32p]),
    "asm_trees_set_device": (_i32, [_p, _i32, _i64, _i32, _p]),
    "asm_rf_block": (_i32, [_p, _i32, _i64, _i64, _i32, _i64, _i64, _i32, _p]),
    "asm_clade_counts": (_i32, [_p, _i32, _i64, _i64, _p]),
    "asm_psrf_sums": (_i32, [_p, _p, _i32, _i32, _i32, _i32, Shard, _p]),
    "asm_psrf_sums_device": (_i32, [_p, _p, _i32, _i32, _i32, _i32, Shard, _p]),
    "asm_psrf_sums_device_async": (_i32, [_p, _p, _i32, _i32, _i32, _i32, Shard, _p]),
    "asm_stream_handle": (_p, [_p]),
    "asm_psrf_finish_device": (_i32, [_p, _p, _i32, _p, _p]),
    "asm_psrf_eval": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "asm_ess_batch": (_i32, [_p, _i32, _i64, _i64, _p, _i32, _p, _p]),
    "asm_ess_batch_device": (_i32, [_p, _i32, _i64, _i64, _p, _i32, _p, _p]),
    "asm_ess_traces_put": (_i32, [_p, _i32, _i64, _i64, _p]),
    "asm_ess_traces_count": (_i32, [_p, _i32p, _i64p]),
    "asm_ess_traces_eval": (_i32, [_p, _i32, _p, _p]),
    "asm_traces_append": (_i32, [_p, _i32, _i32, _i64, _p]),
    "asm_traces_count": (_i32, [_p, _i32, _i64p, _i32p]),
    "asm_ess_eval": (_i32, [_p, _p, _i32, _p, _i32, _p]),
    "asm_trace_moments": (_i32, [_p, _i32, _p, _p]),
    "asm_timer_mark": (_i32, [_p, _i32]),
    "asm_timer_elapsed_ms": (_i32, [_p, _i32, _i32, C.POINTER(C.c_float)]),
    "asm_profile_enable": (_i32, [_p, _i32]),
    "asm_profile_reset": (_i32, [_p]),
    "asm_profile_get": (_i32, [_p, _i32, _f64p, _i64p]),
    "asm_launch_count": (_i64, [_p]),
    "asm_sync": (_i32, [_p]),
    "asm_flush_l2": (_i32, [_p]),
    "asm_ess_last_lag_count": (_i64, [_p]),
    "asm_ess_set_adaptive": (_i32, [_p, _i32]),
    "asm_psrf_last_layout": (_i32, [_p, _i64p, _i64p, _i64p, _i64p]),
    "asm_encoder_create": (_i32, [C.POINTER(_p), _i32]),
    "asm_seq_sum": (_i32, [_p, _p, _i64, _p]),
    "asm_trees_set_ids": (_i32, [_p, _i32, _i64, _i32, _p, _i32]),
    "asm_trees_append_ids": (_i32, [_p, _i32, _i64, _i32, _p, _i32]),
    "asm_bind_thread_near_device": (_i32, [_i32]),
    "asm_encoder_destroy": (_i32, [_p]),
    "asm_encoder_add_topologies": (_i32, [_p, _i64, _p, _p, _p, _p]),
    "asm_encoder_add_newick": (_i32, [_p, C.c_char_p, _i32, _p]),
    "asm_encoder_add_newick_batch": (_i32, [_p, _i64, _p, _i32, _p]),
    "asm_encoder_num_clades": (_i64, [_p]),
    "asm_encoder_clade_leaves": (_i32, [_p, _i64, _p, _i32]),
    "asm_encoder_words": (_i32, [_p]),
    "asm_encode_bits": (_i32, [_p, _i64, _i32, _i32, _p]),
}

_lib: Optional[C.CDLL] = None


def _prefer_bundled_nccl():
    """libasmgpu.so dlopens NCCL by soname on first multi-GPU use.  If this process later imports torch, torch needs
    ITS OWN libnccl.so.2 (a newer one than the system's) under the same soname; point the library at that file now,
    without importing torch, so that whichever loads first the process ends up with one NCCL that suits both."""
    if os.environ.get("ASM_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["ASM_NCCL_LIB"] = cand
                return
    except Exception:  # noqa: BLE001 -- no bundled NCCL: the library falls back to the system's libnccl.so.2
        pass


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Load libasmgpu.so and bind every symbol of the header.  Raises if the library is missing:
    there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `make -C asm_b200/csrc` (or __graft_entry__.build())")
    _prefer_bundled_nccl()
    lib_ = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib_, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib_
    return lib_


def lib() -> C.CDLL:
    return load_library()


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32arr(x: Sequence[int]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


class Encoder:
    """Host encoder: trees -> clade dictionary ids -> bit rows (asm_encoder_* of the C ABI)."""

    def __init__(self, n_taxa: int):
        self._h = C.c_void_p()
        rc = lib().asm_encoder_create(C.byref(self._h), n_taxa)
        if rc != ASM_OK:
            raise AsmGpuError(rc, "asm_encoder_create")
        self.n_taxa = n_taxa

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.asm_encoder_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def add_topologies(self, left: np.ndarray, right: np.ndarray, root: np.ndarray) -> np.ndarray:
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        root = np.ascontiguousarray(root, dtype=np.int32)
        n = root.shape[0]
        ids = np.empty((n, self.n_taxa - 1), dtype=np.int32)
        rc = lib().asm_encoder_add_topologies(self._h, n, _ptr(left), _ptr(right), _ptr(root), _ptr(ids))
        if rc != ASM_OK:
            raise AsmGpuError(rc, "asm_encoder_add_topologies: not a binary tree on n_taxa leaves")
        return ids

    def add_newick(self, newick: str, label_offset: int = 1) -> np.ndarray:
        ids = np.empty(self.n_taxa - 1, dtype=np.int32)
        rc = lib().asm_encoder_add_newick(self._h, newick.encode(), label_offset, _ptr(ids))
        if rc != ASM_OK:
            raise AsmGpuError(rc, "asm_encoder_add_newick")
        return ids

    def add_newick_batch(self, newicks: Sequence[str], label_offset: int = 1) -> np.ndarray:
        n = len(newicks)
        ids = np.empty((n, self.n_taxa - 1), dtype=np.int32)
        arr = (C.c_char_p * max(n, 1))(*[s.encode() for s in newicks])
        rc = lib().asm_encoder_add_newick_batch(self._h, n, arr, label_offset, _ptr(ids))
        if rc != ASM_OK:
            raise AsmGpuError(rc, "asm_encoder_add_newick_batch")
        return ids

    @property
    def num_clades(self) -> int:
        return int(lib().asm_encoder_num_clades(self._h))

    @property
    def words(self) -> int:
        return int(lib().asm_encoder_words(self._h))

    def clade_leaves(self, cid: int) -> np.ndarray:
        buf = np.empty(self.n_taxa, dtype=np.int32)
        n = lib().asm_encoder_clade_leaves(self._h, cid, _ptr(buf), self.n_taxa)
        if n < 0:
            raise AsmGpuError(n, "asm_encoder_clade_leaves")
        return buf[:n].copy()

    def bits(self, clade_ids: np.ndarray, words: Optional[int] = None) -> np.ndarray:
        clade_ids = np.ascontiguousarray(clade_ids, dtype=np.int32)
        if clade_ids.ndim == 1:
            clade_ids = clade_ids[None, :]
        w = self.words if words is None else words
        out = np.empty((clade_ids.shape[0], w), dtype=np.uint64)
        rc = lib().asm_encode_bits(_ptr(clade_ids), clade_ids.shape[0], clade_ids.shape[1], w, _ptr(out))
        if rc != ASM_OK:
            raise AsmGpuError(rc, "asm_encode_bits")
        return out


ASM_COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """asm_comm_unique_id: the bytes rank 0 hands to the other ranks (over the launcher's own channel)."""
    buf = (C.c_uint8 * ASM_COMM_ID_BYTES)()
    rc = lib().asm_comm_unique_id(buf)
    if rc != ASM_OK:
        raise AsmGpuError(rc, (lib().asm_last_error(None) or b"").decode())
    return bytes(buf)


class GpuContext:
    """One asm_ctx: one GPU (`device`) or several GPUs of one box driven from this process (`devices`,
    asm_create_multi).  Mirrors the C ABI one to one; numpy arrays in, numpy arrays out."""

    def __init__(self, n_chains: int, device: int = 0, devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        if devices is not None:
            d = _i32arr(devices)
            rc = lib().asm_create_multi(C.byref(self._h), _ptr(d), d.shape[0], n_chains)
            device = int(d[0]) if d.shape[0] else 0
        else:
            rc = lib().asm_create(C.byref(self._h), device, n_chains)
        if rc != ASM_OK:
            raise AsmGpuError(rc, (lib().asm_last_error(None) or b"").decode())
        self.n_chains = n_chains
        self.device = device
        self.devices = list(devices) if devices is not None else [device]

    # ---- communicator (one process per GPU)
    def comm_init_rank(self, comm_id: bytes, rank: int, n_ranks: int):
        assert len(comm_id) == ASM_COMM_ID_BYTES
        buf = (C.c_uint8 * ASM_COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._check(lib().asm_comm_init_rank(self._h, buf, rank, n_ranks))

    def comm_info(self):
        r, n, l = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(lib().asm_comm_info(self._h, C.byref(r), C.byref(n), C.byref(l)))
        return {"rank": r.value, "n_ranks": n.value, "n_local_devices": l.value}

    def all_gather_f64(self, local: np.ndarray) -> np.ndarray:
        local = np.ascontiguousarray(local, dtype=np.float64)
        n_ranks = self.comm_info()["n_ranks"] if len(self.devices) == 1 else 1
        out = np.empty(local.shape[0] * n_ranks, dtype=np.float64)
        self._check(lib().asm_comm_all_gather_f64(self._h, _ptr(local), local.shape[0], _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.asm_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != ASM_OK:
            raise AsmGpuError(rc, (lib().asm_last_error(self._h) or b"").decode())

    # ---- trees
    def trees_set(self, chain: int, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        self._check(lib().asm_trees_set(self._h, chain, bits.shape[0], bits.shape[1], _ptr(bits)))

    def trees_append(self, chain: int, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        if bits.ndim == 1:
            bits = bits[None, :]
        self._check(lib().asm_trees_append(self._h, chain, bits.shape[0], bits.shape[1], _ptr(bits)))

    @staticmethod
    def _ids16(ids):
        """Clade ids as the ABI takes them: uint16, ASM_NO_CLADE (0xFFFF) in unused slots (the encoder writes -1 there)."""
        ids = np.asarray(ids)
        if ids.dtype != np.uint16:
            if ids.size and (ids.max() > 0xFFFE or ids.min() < -1):
                raise AsmGpuError(ASM_E_RANGE, "clade ids do not fit uint16; use trees_set with bit rows")
            ids = ids.astype(np.uint16)  # -1 -> 0xFFFF
        ids = np.ascontiguousarray(ids)
        return ids[None, :] if ids.ndim == 1 else ids

    def trees_set_ids(self, chain: int, ids, n_clades: int):
        ids = self._ids16(ids)
        self._check(lib().asm_trees_set_ids(self._h, chain, ids.shape[0], ids.shape[1], _ptr(ids), n_clades))

    def trees_append_ids(self, chain: int, ids, n_clades: int):
        ids = self._ids16(ids)
        self._check(lib().asm_trees_append_ids(self._h, chain, ids.shape[0], ids.shape[1], _ptr(ids), n_clades))

    def trees_set_device(self, chain: int, d_ptr: int, n_trees: int, words: int):
        self._check(lib().asm_trees_set_device(self._h, chain, n_trees, words, C.c_void_p(d_ptr)))

    def trees_count(self, chain: int):
        n, w = C.c_int64(), C.c_int32()
        self._check(lib().asm_trees_count(self._h, chain, C.byref(n), C.byref(w)))
        return n.value, w.value

    # ---- probes
    def rf_block(self, chainA: int, r0: int, r1: int, chainB: int, c0: int, c1: int, method: int = ASM_RF_POPC):
        out = np.empty((max(r1 - r0, 0), max(c1 - c0, 0)), dtype=np.uint16)
        self._check(lib().asm_rf_block(self._h, chainA, r0, r1, chainB, c0, c1, method, _ptr(out)))
        return out

    def clade_counts(self, chain: int, start: int, stop: int):
        _, w = self.trees_count(chain)
        out = np.zeros(w * 64, dtype=np.int32)
        self._check(lib().asm_clade_counts(self._h, chain, start, stop, _ptr(out)))
        return out

    # ---- PSRF
    def n_rows(self, row_start: int, end: int, delta: int) -> int:
        return (end - row_start + delta - 1) // delta if end > row_start else 0

    def psrf_sums(self, burnin, row_start: int, end: int, delta: int = 1, method: int = ASM_RF_POPC,
                  shard=(0, 1)) -> np.ndarray:
        b = _i32arr(burnin)
        n = self.n_rows(row_start, end, delta)
        S = np.zeros((self.n_chains, n, self.n_chains), dtype=np.int64)
        self._check(lib().asm_psrf_sums(self._h, _ptr(b), row_start, end, delta, method, Shard(*shard), _ptr(S)))
        return S

    def psrf_sums_device(self, burnin, row_start: int, end: int, delta: int, method: int, shard, d_S_ptr: int):
        b = _i32arr(burnin)
        self._check(lib().asm_psrf_sums_device(self._h, _ptr(b), row_start, end, delta, method, Shard(*shard),
                                               C.c_void_p(d_S_ptr)))

    def psrf_sums_device_async(self, burnin, row_start: int, end: int, delta: int, method: int, shard, d_S_ptr: int):
        """Enqueue only: the caller orders its collective on `stream_handle()`."""
        b = _i32arr(burnin)
        self._check(lib().asm_psrf_sums_device_async(self._h, _ptr(b), row_start, end, delta, method, Shard(*shard),
                                                     C.c_void_p(d_S_ptr)))

    def stream_handle(self) -> int:
        return int(lib().asm_stream_handle(self._h) or 0)

    def psrf_finish_device(self, d_S_ptr: int, n_rows: int, want_rows: bool = False):
        Cn = self.n_chains
        mean = np.empty((Cn, Cn), dtype=np.float64)
        rows = np.empty((Cn, n_rows, Cn), dtype=np.float64) if want_rows else None
        self._check(lib().asm_psrf_finish_device(self._h, C.c_void_p(d_S_ptr), n_rows, _ptr(rows), _ptr(mean)))
        return (rows, mean) if want_rows else mean

    def seq_sum(self, values) -> float:
        """asm_seq_sum: the left-to-right sum of `values`, rounded add by add, evaluated in parallel on the device."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = C.c_double()
        self._check(lib().asm_seq_sum(self._h, _ptr(v), v.shape[0], C.byref(out)))
        return out.value

    def psrf_eval(self, burnin, row_start: int, end: int, delta: int = 1, method: int = ASM_RF_POPC,
                  want_rows: bool = False):
        b = _i32arr(burnin)
        Cn = self.n_chains
        n = self.n_rows(row_start, end, delta)
        mean = np.empty((Cn, Cn), dtype=np.float64)
        rows = np.empty((Cn, n, Cn), dtype=np.float64) if want_rows else None
        self._check(lib().asm_psrf_eval(self._h, _ptr(b), row_start, end, delta, method, _ptr(rows), _ptr(mean)))
        return (rows, mean) if want_rows else mean

    # ---- ESS
    def ess_batch(self, traces: np.ndarray, max_lag: int = ASM_MAX_LAG):
        traces = np.asarray(traces, dtype=np.float64)
        if traces.ndim == 1:
            traces = traces[None, :]
        assert traces.strides[1] == 8 or traces.shape[1] <= 1
        n_tr, n = traces.shape
        stride = traces.strides[0] // 8 if n_tr > 1 else max(n, 1)
        act = np.empty(n_tr, dtype=np.float64)
        ess = np.empty(n_tr, dtype=np.float64)
        self._check(lib().asm_ess_batch(self._h, n_tr, n, max(stride, n), _ptr(traces), max_lag, _ptr(act), _ptr(ess)))
        return act, ess

    def ess_batch_device(self, d_ptr: int, n_traces: int, n_samples: int, stride: int, max_lag: int = ASM_MAX_LAG):
        act = np.empty(n_traces, dtype=np.float64)
        ess = np.empty(n_traces, dtype=np.float64)
        self._check(lib().asm_ess_batch_device(self._h, n_traces, n_samples, stride, C.c_void_p(d_ptr), max_lag,
                                               _ptr(act), _ptr(ess)))
        return act, ess

    def ess_traces_put(self, traces: np.ndarray):
        traces = np.asarray(traces, dtype=np.float64)
        assert traces.ndim == 2 and (traces.strides[1] == 8 or traces.shape[1] <= 1)
        n_tr, n = traces.shape
        stride = traces.strides[0] // 8 if n_tr > 1 else max(n, 1)
        self._check(lib().asm_ess_traces_put(self._h, n_tr, n, max(stride, n), _ptr(traces)))

    def ess_traces_eval(self, max_lag: int = ASM_MAX_LAG):
        n = C.c_int32()
        self._check(lib().asm_ess_traces_count(self._h, C.byref(n), None))
        act = np.empty(n.value, dtype=np.float64)
        ess = np.empty(n.value, dtype=np.float64)
        self._check(lib().asm_ess_traces_eval(self._h, max_lag, _ptr(act), _ptr(ess)))
        return act, ess

    def traces_append(self, chain: int, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[None, :]
        self._check(lib().asm_traces_append(self._h, chain, values.shape[1], values.shape[0], _ptr(values)))

    def traces_count(self, chain: int):
        n, c = C.c_int64(), C.c_int32()
        self._check(lib().asm_traces_count(self._h, chain, C.byref(n), C.byref(c)))
        return n.value, c.value

    def ess_eval(self, burnin, end: int, cols) -> np.ndarray:
        b, cc = _i32arr(burnin), _i32arr(cols)
        out = np.empty(cc.shape[0], dtype=np.float64)
        self._check(lib().asm_ess_eval(self._h, _ptr(b), end, _ptr(cc), cc.shape[0], _ptr(out)))
        return out

    def trace_moments(self, end: int):
        _, n_cols = self.traces_count(0)
        s = np.empty((self.n_chains, n_cols), dtype=np.float64)
        q = np.empty((self.n_chains, n_cols), dtype=np.float64)
        self._check(lib().asm_trace_moments(self._h, end, _ptr(s), _ptr(q)))
        return s, q

    # ---- measurement
    def timer_mark(self, slot: int):
        self._check(lib().asm_timer_mark(self._h, slot))

    def timer_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(lib().asm_timer_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool = True):
        self._check(lib().asm_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(lib().asm_profile_reset(self._h))

    def profile_get(self, family: int):
        ms, n = C.c_double(), C.c_int64()
        self._check(lib().asm_profile_get(self._h, family, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(lib().asm_launch_count(self._h))

    def sync(self):
        self._check(lib().asm_sync(self._h))

    def flush_l2(self):
        self._check(lib().asm_flush_l2(self._h))

    def ess_set_adaptive(self, on: bool):
        self._check(lib().asm_ess_set_adaptive(self._h, int(on)))

    def ess_last_lag_count(self) -> int:
        return int(lib().asm_ess_last_lag_count(self._h))

    def last_layout(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().asm_psrf_last_layout(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"padded_rows": a.value, "padded_bits": b.value, "tiles_done": c.value, "tiles_total": d.value}
