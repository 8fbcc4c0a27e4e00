"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on a CPU-only box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


class SynthChains:
    """Synthetic chains of trees in every representation the tests need: child arrays (for the
    oracle), dictionary ids and bit rows (for the C ABI)."""

    def __init__(self, n_taxa, n_trees, n_chains, beta, moves=2, start_seed=1, seed0=20240001):
        import asm_b200 as A
        from asm_b200 import _lib as L
        import synthgen as G
        import oracle as O

        self.n_taxa, self.n_chains = n_taxa, n_chains
        self.n_trees = [n_trees] * n_chains if np.isscalar(n_trees) else list(n_trees)
        self.enc = A.Encoder(n_taxa)
        self.ts = O.TreeSet(n_chains, n_taxa)
        self.topo, self.ids = [], []
        for c in range(n_chains):
            l, r, root, _ = G.synth_tree_chain(n_taxa, self.n_trees[c], start_seed, seed0 + c, moves, beta)
            self.topo.append((l, r, root))
            self.ids.append(self.enc.add_topologies(l, r, root))
            self.ts.add(c, l, r, root)
        self.words = self.enc.words
        self.bits = [self.enc.bits(i, self.words) for i in self.ids]

    def upload(self, ctx):
        for c in range(self.n_chains):
            ctx.trees_set(c, self.bits[c])


@pytest.fixture(scope="session")
def small_chains():
    """2 chains x 300 trees x 24 taxa: the reference's own regime (two chains)."""
    return SynthChains(24, 300, 2, beta=2.0)


@pytest.fixture(scope="session")
def multi_chains():
    """4 chains x 700 trees x 40 taxa: the generalisation to C > 2 (several 128/256-row tiles per chain)."""
    return SynthChains(40, 700, 4, beta=2.6)


@pytest.fixture(scope="session")
def ragged_chains():
    """3 chains of different lengths."""
    return SynthChains(16, [130, 257, 64], 3, beta=1.5)


def _have_gpu():
    try:
        from asm_b200 import _lib as L
        import synthgen as G
        return L.lib().asm_device_count() > 0
    except Exception:
        return False


@pytest.fixture()
def gpu():
    from asm_b200 import _lib as L
    import synthgen as G
    if L.lib().asm_device_count() == 0:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box (there is no CPU fallback)")
    return L


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-pattern equality of double arrays (NaN == NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.uint64)
    both_nan = np.isnan(a.view(np.float64)) & np.isnan(b.view(np.float64))
    return bool(np.all((a == b) | both_nan))
