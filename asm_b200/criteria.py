"""ctypes binding of include/asmgpu_host.h: the reference's criterion interface (TraceInfo, TreePSRF,
TreeESS, TraceESS; package asm.inference) as implemented in C++ above the C ABI.  Names, inputs and
defaults are the reference's."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib as L

_p, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_i32p = C.POINTER(C.c_int32)

HOST_SIGNATURES = {
    "asmh_last_error": (C.c_char_p, []),
    "asmh_trace_info_create": (_i32, [C.POINTER(_p), _i32, _i32]),
    "asmh_trace_info_destroy": (_i32, [_p]),
    "asmh_read_log_line": (_i32, [_p, _i32, C.c_char_p, _i32p]),
    "asmh_read_tree_log_line": (_i32, [_p, _i32, C.c_char_p, _i32p]),
    "asmh_parse_log_cell": (_f64, [C.c_char_p, _i32p]),
    "asmh_flush": (_i32, [_p]),
    "asmh_set_taxon_count": (_i32, [_p, _i32]),
    "asmh_add_tree": (_i32, [_p, _i32, _p, _p, _i32]),
    "asmh_add_log_line": (_i32, [_p, _i32, _p, _i32]),
    "asmh_set_labels": (_i32, [_p, C.c_char_p]),
    "asmh_get_map": (_i32, [_p, _p, _i32, _i32p]),
    "asmh_tree_count": (_i64, [_p, _i32]),
    "asmh_log_count": (_i64, [_p, _i32]),
    "asmh_clade_count": (_i64, [_p]),
    "asmh_ctx": (_p, [_p]),
    "asmh_tree_psrf_create": (_i32, [C.POINTER(_p), _f64, _i32, _i32, _i32, _f64, _i32, _i32, _i32]),
    "asmh_tree_ess_create": (_i32, [C.POINTER(_p), _i32, _f64, _i32, _i32, _i32]),
    "asmh_trace_ess_create": (_i32, [C.POINTER(_p), _i32, C.c_char_p]),
    "asmh_gelman_rubin_create": (_i32, [C.POINTER(_p), _f64]),
    "asmh_clade_difference_create": (_i32, [C.POINTER(_p), _f64]),
    "asmh_criterion_last_value": (_i32, [_p, C.POINTER(_f64)]),
    "asmh_criterion_destroy": (_i32, [_p]),
    "asmh_criterion_setup": (_i32, [_p, _i32, _p]),
    "asmh_criterion_converged": (_i32, [_p, _p, _i32, _i32p]),
    "asmh_criterion_header": (C.c_char_p, [_p]),
    "asmh_tree_psrf_state": (_i32, [_p, _i32p, _i32p, _i32p, _p, _i32, _i32p]),
    "asmh_trace_ess_current": (_i32, [_p, _p, _i32, _i32p, C.POINTER(_f64)]),
    "asmh_criterion_last_log": (C.c_char_p, [_p]),
    "asmh_burn_in": (_i32, [_p, _i32, _i32, _i32, _p]),
    "asmh_combine_logs": (_i32, [_i32, _i32, C.POINTER(C.c_char_p), _p, _i32, _i64, C.c_char_p, C.c_char_p, C.c_char_p,
                                 C.POINTER(_i64)]),
}

_bound = False


def _h():
    global _bound
    lib = L.lib()
    if not _bound:
        for name, (res, args) in HOST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return lib


class IllegalArgumentException(ValueError):
    pass


def _check(rc: int):
    if rc == L.ASM_OK:
        return
    msg = (_h().asmh_last_error() or b"").decode()
    if rc == L.ASM_E_INVALID:
        raise IllegalArgumentException(msg)
    raise L.AsmGpuError(rc, msg)


def parse_log_cell(token: str):
    """One trace-log cell as AutoStopMCMC.readLogLines converts it (:460-466): (value, parsed) -- value is 0.0 where
    Double.parseDouble would throw.  Host only."""
    ok = C.c_int32()
    v = _h().asmh_parse_log_cell(token.encode(), C.byref(ok))
    return float(v), bool(ok.value)


class TraceInfo:
    """asm.inference.TraceInfo: trees and trace logs of every chain (resident on the GPU)."""

    def __init__(self, chainCount: int, device: int = 0):
        self._p = C.c_void_p()
        _check(_h().asmh_trace_info_create(C.byref(self._p), chainCount, device))
        self._n = chainCount

    def __del__(self):
        if getattr(self, "_p", None) and L._lib is not None:
            L._lib.asmh_trace_info_destroy(self._p)
            self._p = None

    def chainCount(self) -> int:
        return self._n

    def readLogLine(self, chain: int, line: str) -> bool:
        r = C.c_int32()
        _check(_h().asmh_read_log_line(self._p, chain, line.encode(), C.byref(r)))
        return bool(r.value)

    def readTreeLogLine(self, chain: int, line: str) -> bool:
        r = C.c_int32()
        _check(_h().asmh_read_tree_log_line(self._p, chain, line.encode(), C.byref(r)))
        return bool(r.value)

    def setTaxonCount(self, n: int):
        _check(_h().asmh_set_taxon_count(self._p, n))

    def addTree(self, chain: int, left, right, root: int):
        l = np.ascontiguousarray(left, dtype=np.int32)
        r = np.ascontiguousarray(right, dtype=np.int32)
        _check(_h().asmh_add_tree(self._p, chain, l.ctypes.data_as(_p), r.ctypes.data_as(_p), int(root)))

    def addLogLine(self, chain: int, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(_h().asmh_add_log_line(self._p, chain, v.ctypes.data_as(_p), v.shape[0]))

    def setLabels(self, labels: Sequence[str]):
        _check(_h().asmh_set_labels(self._p, " ".join(labels).encode()))

    def getMap(self):
        buf = np.zeros(64, dtype=np.int32)
        n = C.c_int32()
        _check(_h().asmh_get_map(self._p, buf.ctypes.data_as(_p), 64, C.byref(n)))
        return buf[: n.value].tolist()

    def treeCount(self, chain: int) -> int:
        return int(_h().asmh_tree_count(self._p, chain))

    def logCount(self, chain: int) -> int:
        return int(_h().asmh_log_count(self._p, chain))

    def cladeCount(self) -> int:
        return int(_h().asmh_clade_count(self._p))

    def flush(self):
        """Upload what has been appended since the last check (done implicitly by every criterion)."""
        _check(_h().asmh_flush(self._p))


class _Criterion:
    """asm.inference.MCMCConvergenceCriterion"""
    _c: Optional[C.c_void_p] = None

    def __del__(self):
        if getattr(self, "_c", None) and L._lib is not None:
            L._lib.asmh_criterion_destroy(self._c)
            self._c = None

    def setup(self, nChains: int, traceInfo: TraceInfo):
        self._ti = traceInfo
        _check(_h().asmh_criterion_setup(self._c, nChains, traceInfo._p))

    def converged(self, burnin, end: int) -> bool:
        b = np.ascontiguousarray(burnin, dtype=np.int32)
        r = C.c_int32()
        _check(_h().asmh_criterion_converged(self._c, b.ctypes.data_as(_p), end, C.byref(r)))
        return bool(r.value)

    def header(self) -> str:
        return _h().asmh_criterion_header(self._c).decode()

    @property
    def lastLog(self) -> str:
        return _h().asmh_criterion_last_log(self._c).decode()


class TreePSRF(_Criterion):
    """asm.inference.TreePSRF (alias GRLike, version.xml:28).  Inputs: TreePSRF.java:15-20, TreeESS.java:18-26."""

    def __init__(self, b=0.05, twoSided=True, checkESS=False, targetESS=100, smoothing=1.0, cacheLimit=1024,
                 sampleSize=10, method=L.ASM_RF_I8):
        self._c = C.c_void_p()
        _check(_h().asmh_tree_psrf_create(C.byref(self._c), b, int(twoSided), int(checkESS), targetESS, smoothing,
                                          cacheLimit, sampleSize, method))

    def _state(self):
        d, s0, th, n = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        g = np.zeros(64 * 64, dtype=np.float64)
        _check(_h().asmh_tree_psrf_state(self._c, C.byref(d), C.byref(s0), C.byref(th), g.ctypes.data_as(_p), g.shape[0],
                                         C.byref(n)))
        return d.value, s0.value, bool(th.value), g[: n.value].copy()

    delta = property(lambda self: self._state()[0])
    start0 = property(lambda self: self._state()[1])
    threw = property(lambda self: self._state()[2])

    @property
    def grValues(self):
        g = self._state()[3]
        if g.shape[0] == 2:
            return g
        c = int(round(g.shape[0] ** 0.5))
        return g.reshape(c, c)


class TreeESS(_Criterion):
    """asm.inference.TreeESS (its own initAndValidate rejects every smoothing >= 0, TreeESS.java:47)."""

    def __init__(self, targetESS=100, smoothing=1.0, cacheLimit=1024, sampleSize=10, method=L.ASM_RF_I8):
        self._c = C.c_void_p()
        _check(_h().asmh_tree_ess_create(C.byref(self._c), targetESS, smoothing, cacheLimit, sampleSize, method))


class TraceESS(_Criterion):
    """asm.inference.TraceESS.  Inputs: TraceESS.java:14-15."""

    def __init__(self, targetESS=100, traces="posterior,prior,likelihood"):
        self._c = C.c_void_p()
        _check(_h().asmh_trace_ess_create(C.byref(self._c), targetESS, traces.encode()))

    def _cur(self):
        e = np.zeros(64, dtype=np.float64)
        n, mn = C.c_int32(), C.c_double()
        _check(_h().asmh_trace_ess_current(self._c, e.ctypes.data_as(_p), 64, C.byref(n), C.byref(mn)))
        return e[: n.value].copy(), mn.value

    currentESSs = property(lambda self: self._cur()[0])
    minESS = property(lambda self: self._cur()[1])


class GelmanRubin(_Criterion):
    """asm.inference.GelmanRubin.  Input: threshold (GelmanRubin.java:14)."""

    def __init__(self, threshold=1.05):
        self._c = C.c_void_p()
        _check(_h().asmh_gelman_rubin_create(C.byref(self._c), threshold))

    @property
    def lastValue(self) -> float:
        v = C.c_double()
        _check(_h().asmh_criterion_last_value(self._c, C.byref(v)))
        return v.value


class CladeDifference(_Criterion):
    """asm.inference.CladeDifference.  Input: threshold (CladeDifference.java:16)."""

    def __init__(self, threshold=0.25):
        self._c = C.c_void_p()
        _check(_h().asmh_clade_difference_create(C.byref(self._c), threshold))

    @property
    def lastValue(self) -> float:
        v = C.c_double()
        _check(_h().asmh_criterion_last_value(self._c, C.byref(v)))
        return v.value


class BurnInDetector:
    """asm.inference.BurnInDetector (BurnInDetector.java): burnin[] for every converged() call."""

    def __init__(self, traceInfo: "TraceInfo", strategy: str = "Automatic", burnInPercent: Optional[int] = None):
        if strategy not in ("Automatic", "Running"):
            raise IllegalArgumentException("burnInStrat must be Automatic or Running")
        self._ti, self._running = traceInfo, strategy == "Running"
        self._pct = (10 if self._running else 0) if burnInPercent is None else int(burnInPercent)  # BurnInStrategy.java:11-27

    def burnIn(self, end: int):
        out = np.zeros(self._ti.chainCount(), dtype=np.int32)
        _check(_h().asmh_burn_in(self._ti._p, 1 if self._running else 0, self._pct, int(end), out.ctypes.data_as(_p)))
        return out


def combine_logs(tree_mode: bool, chain_paths: Sequence[str], burnin: Sequence[int], last_reported: int,
                 sample_delta: int, out_path: str, header: str = None, footer: str = None) -> int:
    """AutoStopMCMC.combinelogs for one logger (AutoStopMCMC.java:247-317): the samples [burnin[i], last_reported)
    of every chain's log, renumbered 0, sample_delta, ... into out_path.  `header`/`footer` stand for what BEAST's
    Logger prints around them.  Returns the number of samples written; a chain file that ends early raises
    AsmGpuError(ASM_E_RANGE) (the reference's NullPointerException) and leaves out_path alone."""
    n = len(chain_paths)
    paths = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in chain_paths])
    b = np.ascontiguousarray(burnin, dtype=np.int32)
    lines = C.c_int64()
    _check(_h().asmh_combine_logs(1 if tree_mode else 0, n, paths, b.ctypes.data_as(_p), int(last_reported),
                                  int(sample_delta), None if header is None else header.encode(),
                                  None if footer is None else footer.encode(), os.fsencode(out_path), C.byref(lines)))
    return lines.value
