"""ctypes binding of the CPU oracle (oracle/asm_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package.  PARITY UNPINNED -- see the header of asm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libasm_oracle.so")

_p, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_lib: Optional[C.CDLL] = None


class PsrfState(C.Structure):
    """orc_psrf_state: inputs (TreeESS.java:18-26, TreePSRF.java:15-20) + criterion state."""
    _fields_ = [("smoothing", _f64), ("b", _f64), ("targetESS", _i32), ("cacheLimit", _i32), ("sampleSize", _i32),
                ("twoSided", _i32), ("checkESS", _i32), ("delta", _i32), ("start0", _i32), ("prev", _i32),
                ("n_chains", _i32), ("grValues", _f64 * (64 * 64)), ("n_rows_last", _i32), ("threw", _i32),
                ("have_indices", _i32), ("indices", _i32 * 1024)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "asm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_treeset_new.restype = _p
        L.orc_treeset_new.argtypes = [_i32, _i32]
        L.orc_treeset_free.argtypes = [_p]
        L.orc_treeset_add.restype = _i32
        L.orc_treeset_add.argtypes = [_p, _i32, _i64, _p, _p, _p]
        L.orc_num_clades.restype = _i64
        L.orc_num_clades.argtypes = [_p]
        L.orc_num_trees.restype = _i64
        L.orc_num_trees.argtypes = [_p, _i32]
        L.orc_clade_key.restype = _i32
        L.orc_clade_key.argtypes = [_p, _i64, C.c_char_p, _i64]
        L.orc_clade_size.restype = _i32
        L.orc_clade_size.argtypes = [_p, _i64]
        L.orc_tree_clade_ids.argtypes = [_p, _i32, _i64, _p]
        L.orc_clade_counts.argtypes = [_p, _i32, _i64, _i64, _p]
        L.orc_rf.restype = _i32
        L.orc_rf.argtypes = [_p, _i32, _i64, _i32, _i64]
        L.orc_distance_plus_one.restype = C.c_float
        L.orc_distance_plus_one.argtypes = [_p, _i32, _i64, _i32, _i64]
        L.orc_rf_block.argtypes = [_p, _i32, _i64, _i64, _i32, _i64, _i64, _p]
        L.orc_calc_psrf.restype = _f64
        L.orc_calc_psrf.argtypes = [_p, _i32, _i32, _i32, _p, _i32, _i32, _p]
        L.orc_psrf_rows_mean.restype = _f64
        L.orc_psrf_rows_mean.argtypes = [_p, _i32, _i32, _i32, _p, _i32, _i32, _p, _p]
        L.orc_psrf_init.restype = _i32
        L.orc_psrf_init.argtypes = [C.POINTER(PsrfState)]
        L.orc_psrf_setup.restype = _i32
        L.orc_psrf_setup.argtypes = [C.POINTER(PsrfState), _i32]
        L.orc_psrf_converged.restype = _i32
        L.orc_psrf_converged.argtypes = [C.POINTER(PsrfState), _p, _p, _i32]
        L.orc_psrf_sums.argtypes = [_p, _p, _i32, _i32, _i32, _p]
        L.orc_act_literal.restype = _f64
        L.orc_act_literal.argtypes = [_p, _i64, _i64]
        L.orc_act_final.restype = _f64
        L.orc_act_final.argtypes = [_p, _i64, _i64, _i32, _p, _p]
        L.orc_calc_ess.restype = _f64
        L.orc_calc_ess.argtypes = [_p, _i64, _i64, _i32]
        L.orc_ess_batch.argtypes = [_p, _i32, _i64, _i64, _i32, _p, _p]
        L.orc_trace_ess_converged.restype = _i32
        L.orc_trace_ess_converged.argtypes = [_p, _i64, _i32, _p, _i32, _p, _i32, _i32, _i32, _p, _p]
        L.orc_burnin_trace.restype = _i32
        L.orc_burnin_trace.argtypes = [_p, _i32]
        L.orc_burnin.argtypes = [_p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p]
        L.orc_pseudo_ess.restype = _f64
        L.orc_pseudo_ess.argtypes = [_p, _p, _p, _i32, _i32, _i32, _i32, _i32]
        L.orc_gr_stat.restype = _f64
        L.orc_gr_stat.argtypes = [_i32, _p, _p]
        L.orc_gelman_rubin_converged.restype = _i32
        L.orc_gelman_rubin_converged.argtypes = [_p, _i64, _i32, _i32, _i32, _f64, _p]
        L.orc_max_clade_difference.restype = _f64
        L.orc_max_clade_difference.argtypes = [_p, _i32, _i32, _i32]
        L.orc_clade_difference_converged.restype = _i32
        L.orc_clade_difference_converged.argtypes = [_p, _i32, _i32, _f64, C.POINTER(_f64)]
        L.orc_num_threads.restype = _i32
        L.orc_set_num_threads.argtypes = [_i32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


class TreeSet:
    """Oracle stand-in for TraceInfo.trees (TraceInfo.java:27): per-chain lists of trees."""

    def __init__(self, n_chains: int, n_taxa: int):
        self._h = lib().orc_treeset_new(n_chains, n_taxa)
        self.n_chains, self.n_taxa = n_chains, n_taxa

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_treeset_free(self._h)
            self._h = None

    def add(self, chain: int, left: np.ndarray, right: np.ndarray, root: np.ndarray):
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        root = np.ascontiguousarray(root, dtype=np.int32)
        rc = lib().orc_treeset_add(self._h, chain, root.shape[0], _ptr(left), _ptr(right), _ptr(root))
        if rc != 0:
            raise ValueError("not a binary tree on n_taxa leaves")

    @property
    def num_clades(self) -> int:
        return int(lib().orc_num_clades(self._h))

    def num_trees(self, chain: int) -> int:
        return int(lib().orc_num_trees(self._h, chain))

    def clade_key(self, cid: int) -> str:
        buf = C.create_string_buffer(16 * self.n_taxa + 16)
        n = lib().orc_clade_key(self._h, cid, buf, len(buf))
        assert n >= 0
        return buf.value.decode()

    def tree_clade_ids(self, chain: int, idx: int) -> np.ndarray:
        out = np.empty(self.n_taxa - 1, dtype=np.int32)
        lib().orc_tree_clade_ids(self._h, chain, idx, _ptr(out))
        return out

    def clade_counts(self, chain: int, start: int, stop: int) -> np.ndarray:
        out = np.zeros(self.num_clades, dtype=np.int32)
        lib().orc_clade_counts(self._h, chain, start, stop, _ptr(out))
        return out

    def rf(self, c1: int, i1: int, c2: int, i2: int) -> int:
        return int(lib().orc_rf(self._h, c1, i1, c2, i2))

    def distance_plus_one(self, c1, i1, c2, i2) -> float:
        return float(lib().orc_distance_plus_one(self._h, c1, i1, c2, i2))

    def rf_block(self, chainA, r0, r1, chainB, c0, c1) -> np.ndarray:
        out = np.empty((r1 - r0, c1 - c0), dtype=np.uint16)
        lib().orc_rf_block(self._h, chainA, r0, r1, chainB, c0, c1, _ptr(out))
        return out

    def calc_psrf(self, ts1, ts2, k, burnin, end, delta=1):
        b = _i32arr(burnin)
        v = np.empty(2, dtype=np.float64)
        r = lib().orc_calc_psrf(self._h, ts1, ts2, k, _ptr(b), end, delta, _ptr(v))
        return float(r), v

    def psrf_rows_mean(self, ts1, ts2, start, burnin, end, delta=1, want_rows=False):
        b = _i32arr(burnin)
        n = max((end - start) // delta, 0)
        rows = np.empty(n, dtype=np.float64) if want_rows else None
        m = lib().orc_psrf_rows_mean(self._h, ts1, ts2, start, _ptr(b), end, delta, _ptr(rows), None)
        return (float(m), rows) if want_rows else float(m)

    def psrf_sums(self, burnin, start, end, delta=1) -> np.ndarray:
        b = _i32arr(burnin)
        n = (end - start + delta - 1) // delta if end > start else 0
        S = np.zeros((self.n_chains, n, self.n_chains), dtype=np.float64)
        lib().orc_psrf_sums(self._h, _ptr(b), start, end, delta, _ptr(S))
        return S


class TreePSRF:
    """Oracle TreePSRF criterion (TreePSRF.java): inputs by the reference's names, converged(burnin, end)."""

    def __init__(self, b=0.05, twoSided=True, checkESS=False, targetESS=100, smoothing=1.0, cacheLimit=1024,
                 sampleSize=10):
        self.st = PsrfState()
        self.st.smoothing, self.st.b = smoothing, b
        self.st.targetESS, self.st.cacheLimit, self.st.sampleSize = targetESS, cacheLimit, sampleSize
        self.st.twoSided, self.st.checkESS = int(twoSided), int(checkESS)
        rc = lib().orc_psrf_init(C.byref(self.st))  # initAndValidate, TreePSRF.java:54-93
        if rc != 0:
            raise ValueError(f"IllegalArgumentException site {rc}")
        self.ts = None

    def setup(self, nChains: int, treeset: TreeSet):
        rc = lib().orc_psrf_setup(C.byref(self.st), nChains)
        if rc < 0:
            raise ValueError("bad chain count")
        self.ts = treeset

    def converged(self, burnin, end: int) -> bool:
        b = _i32arr(burnin)
        return bool(lib().orc_psrf_converged(C.byref(self.st), self.ts._h, _ptr(b), end))

    @property
    def grValues(self):
        n = self.st.n_chains
        g = np.array(self.st.grValues[: n * n], dtype=np.float64)
        return g[:2].copy() if n == 2 else g.reshape(n, n)

    @property
    def delta(self):
        return self.st.delta

    @property
    def start0(self):
        return self.st.start0


def act_literal(x: np.ndarray, start=0, end=None) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(lib().orc_act_literal(_ptr(x), start, len(x) if end is None else end))


def act_final(x: np.ndarray, start=0, end=None, max_lag=2000, want_ac=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    end = len(x) if end is None else end
    L = min(end - start, max_lag)
    ac = np.empty(max(L, 1), dtype=np.float64) if want_ac else None
    used = C.c_int32()
    r = float(lib().orc_act_final(_ptr(x), start, end, max_lag, _ptr(ac), C.byref(used)))
    return (r, ac[:L], used.value) if want_ac else r


def ess_batch(traces: np.ndarray, literal=False):
    traces = np.ascontiguousarray(traces, dtype=np.float64)
    n_tr, n = traces.shape
    act = np.empty(n_tr)
    ess = np.empty(n_tr)
    lib().orc_ess_batch(_ptr(traces), n_tr, n, n, int(literal), _ptr(act), _ptr(ess))
    return act, ess


def _log_ptrs(logs):
    """logs: list over chains of arrays [n_cols][n_samples] (column-major tables, C-contiguous)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in logs]
    stride = arrs[0].shape[1]
    assert all(a.shape[1] == stride for a in arrs)
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs, stride


def trace_ess_converged(logs, burnin, end, cols, targetESS=100, literal=False):
    """TraceESS.converged (TraceESS.java:38-69). Returns (converged, currentESSs, minESS)."""
    arrs, ptrs, stride = _log_ptrs(logs)
    b, m = _i32arr(burnin), _i32arr(cols)
    ess = np.empty(len(m))
    mn = C.c_double()
    r = lib().orc_trace_ess_converged(ptrs, stride, len(arrs), _ptr(b), end, _ptr(m), len(m), targetESS, int(literal),
                                      _ptr(ess), C.byref(mn))
    return bool(r), ess, mn.value


def burnin(logs, cols, strategy: str, percent: int, end: int) -> np.ndarray:
    """BurnInDetector.burnIn(end) (BurnInDetector.java:21-36)."""
    arrs, ptrs, stride = _log_ptrs(logs)
    m = _i32arr(cols)
    out = np.empty(len(arrs), dtype=np.int32)
    lib().orc_burnin(ptrs, stride, len(arrs), _ptr(m), len(m), 0 if strategy == "Automatic" else 1, percent, end,
                     _ptr(out))
    return out


def gelman_rubin_converged(logs, available: int, threshold: float = 1.05):
    """GelmanRubin.converged (GelmanRubin.java:32-51). Returns (converged, GR statistics [pairs][items])."""
    arrs, ptrs, stride = _log_ptrs(logs)
    n, items = len(arrs), arrs[0].shape[0]
    gr = np.empty((n * (n - 1) // 2, items))
    r = lib().orc_gelman_rubin_converged(ptrs, stride, n, items, available, threshold, _ptr(gr))
    return bool(r), gr


def clade_difference_converged(ts: "TreeSet", available: int, threshold: float = 0.25):
    """CladeDifference.converged (CladeDifference.java:41-65). Returns (converged, fMaxCladeProbDiff)."""
    mx = C.c_double()
    r = lib().orc_clade_difference_converged(ts._h, ts.n_chains, available, threshold, C.byref(mx))
    return bool(r), mx.value


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """Make the OpenMP loops use every core this process may run on, whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its children).  Returns the thread count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(max(n, 1))
    return num_threads()


def combine_logs(tree_mode: bool, chain_paths, burnin, last_reported: int, sample_delta: int) -> str:
    """AutoStopMCMC.combinelogs (AutoStopMCMC.java:247-317) for one logger, restated line by line with Java's
    regular expressions; returns what `out` receives after logger.init().  Raises TypeError where the Java
    dereferences the null of a finished stream (str.matches / str.replaceFirst, :272 / :306-308).
    Test infrastructure only (pure Python: the files of the parity cases are small)."""
    import io
    import re

    class Reader:  # BufferedReader over a file: readLine() splits on \n, \r, \r\n and returns None at the end
        def __init__(self, path):
            data = open(path, "rb").read().decode("latin-1")
            self.lines = re.split(r"\r\n|\n|\r", data)
            if self.lines and self.lines[-1] == "":
                self.lines.pop()
            self.pos = 0

        def ready(self):
            return self.pos < len(self.lines)

        def read_line(self):
            if self.pos >= len(self.lines):
                return None
            self.pos += 1
            return self.lines[self.pos - 1]

    out = io.StringIO()
    sample_nr = 0  # :252
    for i, path in enumerate(chain_paths):
        fin = Reader(path)
        s = None
        if tree_mode:  # :260-285
            skipped = 0
            while fin.ready() and skipped < burnin[i]:
                s = fin.read_line()
                if re.fullmatch(r"^tree STATE.*", s):
                    skipped += 1
            for _ in range(burnin[i], last_reported):
                s = fin.read_line()
                if re.fullmatch(r"^tree STATE.*", s):  # TypeError on None = NullPointerException
                    s = re.sub(r"^tree STATE_[^\s]*", "", s)
                    out.write("tree STATE_" + str(sample_nr) + s + "\n")
                    sample_nr += sample_delta
        else:  # :287-312
            while fin.ready() and (s is None or not s.startswith("Sample")):
                s = fin.read_line()
            skipped = 0
            while fin.ready() and skipped < burnin[i]:
                s = fin.read_line()
                skipped += 1
            for _ in range(burnin[i], last_reported):
                s = fin.read_line()
                s = re.sub(r"[0-9]+\t", "", s, count=1)  # TypeError on None = NullPointerException
                out.write(str(sample_nr) + "\t" + s + "\n")
                sample_nr += sample_delta
    return out.getvalue()
