"""synthgen -- synthetic inputs of the shapes BASELINE.json names (tree chains, AR(1) traces).

Test and benchmark infrastructure, NOT part of the product: `libasmsynth.so` holds the generators and a private
copy of the host encoder, so the reference arm of bench.py (and anything else that only needs inputs) maps nothing
of `asm_b200/libasmgpu.so`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasmsynth.so")
_p, _i32, _i64, _u64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(LIB_PATH)
        L.asm_synth_tree_chain.restype = _i32
        L.asm_synth_tree_chain.argtypes = [_i32, _i64, _u64, _u64, _i32, _f64, _p, _p, _p, _p, _p]
        L.asm_synth_ar1.restype = _i32
        L.asm_synth_ar1.argtypes = [_p, _i64, _f64, _f64, _u64]
        L.asm_encoder_create.restype = _i32
        L.asm_encoder_create.argtypes = [C.POINTER(_p), _i32]
        L.asm_encoder_destroy.argtypes = [_p]
        L.asm_encoder_num_clades.restype = _i64
        L.asm_encoder_num_clades.argtypes = [_p]
        L.asm_encoder_words.restype = _i32
        L.asm_encoder_words.argtypes = [_p]
        L.asm_encode_bits.restype = _i32
        L.asm_encode_bits.argtypes = [_p, _i64, _i32, _i32, _p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Encoder:
    """The generator library's own clade dictionary (same code as the product's host encoder, private copy)."""

    def __init__(self, n_taxa: int):
        self._h = C.c_void_p()
        if lib().asm_encoder_create(C.byref(self._h), n_taxa) != 0:
            raise RuntimeError("asm_encoder_create")
        self.n_taxa = n_taxa

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.asm_encoder_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    @property
    def num_clades(self) -> int:
        return int(lib().asm_encoder_num_clades(self._h))

    @property
    def words(self) -> int:
        return int(lib().asm_encoder_words(self._h))

    def bits(self, clade_ids: np.ndarray, words: Optional[int] = None) -> np.ndarray:
        clade_ids = np.ascontiguousarray(clade_ids, dtype=np.int32)
        if clade_ids.ndim == 1:
            clade_ids = clade_ids[None, :]
        w = self.words if words is None else words
        out = np.empty((clade_ids.shape[0], w), dtype=np.uint64)
        if lib().asm_encode_bits(_ptr(clade_ids), clade_ids.shape[0], clade_ids.shape[1], w, _ptr(out)) != 0:
            raise RuntimeError("asm_encode_bits")
        return out


def synth_tree_chain(n_taxa: int, n_trees: int, start_seed: int, chain_seed: int, moves_per_sample: int = 2,
                     beta: float = 7.0, topologies: bool = True, encoder: Optional[Encoder] = None):
    """asm_synth_tree_chain: returns (left, right, root, clade_ids); entries not requested are None."""
    if encoder is not None and not isinstance(encoder, Encoder):
        raise TypeError("encoder must be a synthgen.Encoder (the generator library's own dictionary)")
    nn = 2 * n_taxa - 1
    left = np.empty((n_trees, nn), dtype=np.int32) if topologies else None
    right = np.empty((n_trees, nn), dtype=np.int32) if topologies else None
    root = np.empty(n_trees, dtype=np.int32) if topologies else None
    ids = np.empty((n_trees, n_taxa - 1), dtype=np.int32) if encoder is not None else None
    rc = lib().asm_synth_tree_chain(n_taxa, n_trees, start_seed, chain_seed, moves_per_sample, beta, _ptr(left),
                                    _ptr(right), _ptr(root), encoder.handle if encoder is not None else None, _ptr(ids))
    if rc != 0:
        raise RuntimeError(f"asm_synth_tree_chain failed ({rc})")
    return left, right, root, ids


def synth_ar1(n: int, mu: float, phi: float, seed: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    if out is None:
        out = np.empty(n, dtype=np.float64)
    assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape[0] >= n
    if lib().asm_synth_ar1(_ptr(out), n, mu, phi, seed) != 0:
        raise RuntimeError("asm_synth_ar1")
    return out
