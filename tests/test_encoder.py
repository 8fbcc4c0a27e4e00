"""Host encoder (the step before the path): Newick / child arrays -> clade ids -> bit rows."""
import numpy as np
import pytest

import asm_b200 as A
import oracle as O
from asm_b200 import _lib as L
import synthgen as G


def test_ids_follow_the_reference_traverse_order(multi_chains):
    ch = multi_chains
    # same post-order, same first-seen numbering as CladeDifference.traverse in the oracle
    for c in range(ch.n_chains):
        for i in (0, 1, 350, 699):
            assert np.array_equal(ch.ids[c][i], ch.ts.tree_clade_ids(c, i))
    assert ch.enc.num_clades == ch.ts.num_clades
    for cid in range(0, ch.enc.num_clades, 37):
        assert "".join(f"{x}," for x in ch.enc.clade_leaves(cid)) == ch.ts.clade_key(cid)


def test_bit_rows(multi_chains):
    ch = multi_chains
    assert ch.words % 4 == 0 and ch.words * 64 >= ch.enc.num_clades
    row = ch.bits[2][17]
    ids = sorted(int(w) * 64 + b for w in range(ch.words) for b in range(64) if (int(row[w]) >> b) & 1)
    assert ids == sorted(ch.ids[2][17].tolist())
    with pytest.raises(L.AsmGpuError):
        ch.enc.bits(ch.ids[3][-2:], words=1)  # dictionary does not fit


def test_newick_reader():
    enc = A.Encoder(5)
    # BEAST tree log style: translate indices start at 1 (parser.offsetInput = 1, AutoStopMCMC.java:527)
    a = enc.add_newick("((1:0.1,2:0.2):0.3,(3:0.1,(4:0.05,5:0.05):0.2):0.1):0.0;", 1)
    b = enc.add_newick("((1[&rate=1.0]:0.1,2:0.2)[&posterior=0.9]:0.3,(3:0.1,(4:0.05,5:0.05):0.2):0.1);", 1)
    assert np.array_equal(a, b)
    leaves = [enc.clade_leaves(int(i)).tolist() for i in a]
    assert leaves == [[0, 1], [3, 4], [2, 3, 4], [0, 1, 2, 3, 4]]  # post-order, left before right
    c = enc.add_newick("((2,1),((5,4),3));", 1)                     # same clades, different child order
    assert sorted(c.tolist()) == sorted(a.tolist())
    for bad in ("((1,2),(3,4));", "((1,2),(3,(4,5)),6);", "((1,1),(3,(4,5)));", "((a,b),(c,(d,e)));", "((1,2,3),(4,5));"):
        with pytest.raises(L.AsmGpuError):
            enc.add_newick(bad, 1)


def test_topology_validation():
    enc = A.Encoder(4)
    left = np.array([[-1, -1, -1, -1, 0, 2, 4]], dtype=np.int32)
    right = np.array([[-1, -1, -1, -1, 1, 3, 5]], dtype=np.int32)
    ids = enc.add_topologies(left, right, np.array([6], dtype=np.int32))
    assert [enc.clade_leaves(int(i)).tolist() for i in ids[0]] == [[0, 1], [2, 3], [0, 1, 2, 3]]
    bad = right.copy()
    bad[0, 6] = 4  # both children the same subtree: not a tree on 4 leaves
    with pytest.raises(L.AsmGpuError):
        enc.add_topologies(left, bad, np.array([6], dtype=np.int32))


def test_synthetic_generators_are_reproducible():
    a = G.synth_tree_chain(30, 50, 1, 42, 2, 2.5)
    b = G.synth_tree_chain(30, 50, 1, 42, 2, 2.5)
    c = G.synth_tree_chain(30, 50, 1, 43, 2, 2.5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], c[0])
    # the fast path (ids straight from the walk) names the same clades as the canonical encoder
    enc1, enc2 = G.Encoder(30), A.Encoder(30)  # the generator library's private dictionary vs the product's encoder
    _, _, _, ids_fast = G.synth_tree_chain(30, 50, 1, 42, 2, 2.5, topologies=False, encoder=enc1)
    ids_ref = enc2.add_topologies(*a[:3])
    for t in (0, 10, 49):
        assert np.array_equal(np.bitwise_count(enc1.bits(ids_fast[t])), np.bitwise_count(enc2.bits(ids_ref[t])))
    # same pairwise distances from either dictionary (clade numbering differs, the sets of clades do not)
    bf, br = enc1.bits(ids_fast), enc2.bits(ids_ref)
    df = np.bitwise_count(bf[:, None, :] ^ bf[None, :, :]).sum(-1)
    dr = np.bitwise_count(br[:, None, :] ^ br[None, :, :]).sum(-1)
    assert np.array_equal(df, dr)
    x = G.synth_ar1(1000, -1e4, 0.9, 5)
    assert np.array_equal(x, G.synth_ar1(1000, -1e4, 0.9, 5)) and abs(x.mean() + 1e4) < 5


def _newick_of(left, right, root, offset=1):
    """Child arrays -> Newick with integer labels (leaf number + offset), iteratively."""
    out, stack = [], [(int(root), 0)]
    while stack:
        v, stage = stack.pop()
        if left[v] < 0:
            out.append(str(v + offset))
        elif stage == 0:
            out.append("(")
            stack.append((v, 1))
            stack.append((int(left[v]), 0))
        elif stage == 1:
            out.append(",")
            stack.append((v, 2))
            stack.append((int(right[v]), 0))
        else:
            out.append(")")
    return "".join(out) + ";"


@pytest.mark.parametrize("threads", [2, 5, 8])
def test_parallel_encoder_hands_out_the_serial_ids(monkeypatch, threads):
    # two-phase batch encoder (look-ups against the frozen dictionary on every core, insertions in first-seen order):
    # the ids are those of the one-tree-at-a-time encoder, i.e. CladeDifference.traverse's numbering
    n_taxa, n_trees = 60, 3000
    l, r, root, _ = G.synth_tree_chain(n_taxa, n_trees, 1, 777, 2, 3.0)
    l2, r2, root2, _ = G.synth_tree_chain(n_taxa, 700, 1, 778, 2, 3.0)
    monkeypatch.setenv("ASM_ENCODER_THREADS", "1")
    serial = A.Encoder(n_taxa)
    ids_s = serial.add_topologies(l, r, root)
    ids_s2 = serial.add_topologies(l2, r2, root2)
    monkeypatch.setenv("ASM_ENCODER_THREADS", str(threads))
    par = A.Encoder(n_taxa)
    ids_p = par.add_topologies(l, r, root)
    ids_p2 = par.add_topologies(l2, r2, root2)  # second call: the dictionary is warm, most clades are hits
    assert np.array_equal(ids_s, ids_p) and np.array_equal(ids_s2, ids_p2)
    assert serial.num_clades == par.num_clades
    for cid in range(0, serial.num_clades, 101):
        assert np.array_equal(serial.clade_leaves(cid), par.clade_leaves(cid))
    assert np.array_equal(serial.bits(ids_s), par.bits(ids_p))  # asm_encode_bits is threaded the same way


def test_newick_batch_matches_one_by_one(monkeypatch):
    n_taxa, n_trees = 33, 1500
    l, r, root, _ = G.synth_tree_chain(n_taxa, n_trees, 1, 4242, 2, 2.5)
    newicks = [_newick_of(l[t], r[t], root[t]) for t in range(n_trees)]
    one = A.Encoder(n_taxa)
    ids_one = np.stack([one.add_newick(s, 1) for s in newicks])
    monkeypatch.setenv("ASM_ENCODER_THREADS", "6")
    batch = A.Encoder(n_taxa)
    ids_batch = batch.add_newick_batch(newicks, 1)
    assert np.array_equal(ids_one, ids_batch)
    # and both equal the child-array path
    topo = A.Encoder(n_taxa)
    assert np.array_equal(topo.add_topologies(l, r, root), ids_batch)
    # a malformed tree anywhere in a batch is reported with the reader's own error code
    bad = list(newicks)
    bad[1234] = "((1,2,3),(4,5));"
    with pytest.raises(L.AsmGpuError) as ei:
        A.Encoder(n_taxa).add_newick_batch(bad, 1)
    assert ei.value.code in (L.ASM_E_UNSUPPORTED, L.ASM_E_INVALID)
    assert A.Encoder(n_taxa).add_newick_batch([], 1).shape == (0, n_taxa - 1)
