"""GPU parity for the per-clade scan (st_clade_moments / SuchLinkedTrees.clade_pearson):
the reference's loop `for node: subset_b(node); linked_distances(); pearson()`
(docs/examples/SuchLinkedTree_examples.md:286-310) as one launch sequence."""
import os

import numpy as np
import pandas as pd
import pytest
from conftest import DATA, GOLDEN

import oracle as O
from suchtree_b200 import SuchLinkedTrees, SuchTree, as_moments, moments_pearson, synth

pytestmark = pytest.mark.gpu

LINKED = {
    "gopher_louse": ("test.tree", "lice.tree", "links.csv"),
    "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"),
    "perfect0": ("perfect0_host.tree", "perfect0_guest.tree", "perfect0_links.csv"),
    "arr1": ("arr1_plant.tree", "arr1_animal.tree", "arr1_links.csv"),
}


def _slt(name):
    ta, tb, lk = LINKED[name]
    T1 = SuchTree(os.path.join(DATA, ta))
    T2 = SuchTree(os.path.join(DATA, tb))
    links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
    if set(links.index) != set(T1.leaves.keys()):
        links = links.T
    return SuchLinkedTrees(T1, T2, links), T1, T2


def _oracle(T):
    return O.OracleTree(T._ft.parent, T._ft.distance, depth=T.depth)


@pytest.mark.parametrize("scan", ["b", "a", "ab"])
@pytest.mark.parametrize("name", list(LINKED))
def test_clade_scan_matches_reference_loop(name, scan):
    """Against the reference's own loop (tests/golden/make_golden_clades.py): subset sizes and
    link counts exact; r within the reference's fp32 path accumulation of its fp64 correlation,
    and within 1e-9 of the oracle's restatement of the loop (O2 distances, fp64 pearson)."""
    z = np.load(os.path.join(GOLDEN, "clades.npz"))
    g = {k: z["%s__%s__%s" % (name, scan, k)] for k in ("nodes", "n_leafs", "n_links", "r32", "r64")}
    SLT, T1, T2 = _slt(name)
    side = "a" if scan == "a" else "b"
    if scan == "ab":
        SLT.subset_a(int(z[name + "__ab__anode"]))
    ll = SLT.linklist.copy()
    res = SLT.clade_pearson(g["nodes"], side=side)
    assert np.array_equal(res["node_ids"], g["nodes"])
    assert np.array_equal(res["n_leafs"], g["n_leafs"])
    assert np.array_equal(res["n_links"], g["n_links"])
    assert np.array_equal(res["n_pairs"], g["n_links"] * (g["n_links"] - 1) // 2 * (g["n_links"] >= 2))
    assert np.all(np.isnan(res["r"][g["n_links"] < 2]))
    ok = np.isfinite(g["r64"])
    assert np.allclose(res["r"][ok], g["r64"][ok], rtol=0, atol=5e-6)
    T = T1 if side == "a" else T2
    n_links, r = O.clade_scan(_oracle(T1), _oracle(T2), T._ft.left, T._ft.right, ll, g["nodes"], side)
    assert np.array_equal(res["n_links"], n_links)
    has = n_links >= 2
    assert np.allclose(res["r"][has], r[has], rtol=0, atol=1e-9)
    # the scan leaves the object's own subset state alone
    assert np.array_equal(SLT.linklist, ll)


@pytest.mark.parametrize("name", ["fishworm", "arr1"])
def test_clade_scan_equals_the_products_own_loop(name):
    """subset_b(node) + linked_pearson() per node (one host round trip per clade) gives the
    same r; default nodes = all internal nodes of the scanned tree; min/max_links skip."""
    SLT, T1, T2 = _slt(name)
    res = SLT.clade_pearson(min_links=5, max_links=60)
    assert np.array_equal(res["node_ids"], np.nonzero(T2._ft.left != -1)[0])
    skipped = (res["n_links"] < 5) | (res["n_links"] > 60)
    assert skipped.any() and (~skipped).any()
    assert np.all(np.isnan(res["r"][skipped])) and not res["n_pairs"][skipped].any()
    for node, n_links, r in zip(res["node_ids"], res["n_links"], res["r"]):
        SLT.subset_b(int(node))
        assert SLT.subset_n_links == n_links
        if 5 <= n_links <= 60:
            assert r == pytest.approx(SLT.linked_pearson(), abs=1e-10)
    SLT.subset_b(T2.root_node)
    # moments of the root clade = the exhaustive moments of the whole link list
    _, _, nl, mom = SLT.clade_moments([T2.root_node])
    assert nl[0] == SLT.n_links and mom.shape == (1, 8)
    m0 = as_moments(mom[0])
    assert moments_pearson(m0) == pytest.approx(SLT.linked_pearson(), abs=1e-12)
    whole = SLT.linked_moments(x0=m0.x0, y0=m0.y0)
    for k in ("n", "sx", "sy", "sxx", "syy", "sxy"):
        assert getattr(m0, k) == pytest.approx(getattr(whole, k), rel=1e-11, abs=1e-9)


@pytest.mark.parametrize("wide", [False, True])
def test_clade_scan_synthetic_many_items(wide):
    """2000-leaf Yule trees, 3000 random links (several links per leaf, unlinked leaves):
    clades from 2 links to all of them, so work items range from one pair to many 4096-pair
    chunks of one clade; against the oracle's restatement of the loop.  Both index layouts."""
    fa, fb = synth.yule_tree(2000, seed=14), synth.yule_tree(2000, seed=15)
    TA, TB = SuchTree.from_flat(fa, _wide=wide), SuchTree.from_flat(fb, _wide=wide)
    rng = np.random.default_rng(16)
    L = 3000
    ll = np.stack([2 * rng.integers(0, 2000, L), 2 * rng.integers(0, 2000, L)], axis=1).astype(np.int64)
    SLT = SuchLinkedTrees.from_linklist(TA, TB, ll)
    assert SLT.n_links == L and np.array_equal(np.sort(SLT.linklist[:, 0]), SLT.linklist[:, 0])
    oa, ob = O.OracleTree(fa.parent, fa.distance), O.OracleTree(fb.parent, fb.distance)
    for side, T in (("b", TB), ("a", TA)):
        res = SLT.clade_pearson(side=side)
        inner = np.nonzero(T._ft.left != -1)[0]
        assert np.array_equal(res["node_ids"], inner) and res["n_links"].max() == L
        lo, hi = T._clade_intervals()
        col = SLT.linklist[:, 0 if side == "b" else 1]
        want_links = np.array([((col >= lo[v]) & (col <= hi[v])).sum() for v in inner])
        assert np.array_equal(res["n_links"], want_links)
        # oracle loop on a spread of clade sizes (the root included)
        order = np.argsort(res["n_links"])
        pick = np.unique(np.concatenate([order[:: max(1, len(order) // 40)], order[-3:]]))
        n_links, r = O.clade_scan(oa, ob, T._ft.left, T._ft.right, SLT.linklist, inner[pick], side)
        assert np.array_equal(n_links, res["n_links"][pick])
        has = n_links >= 2
        assert has.sum() >= 10
        assert np.allclose(res["r"][pick][has], r[has], rtol=0, atol=1e-9)
    # run to run: identical bits (per-item partials folded in a fixed order)
    again = SLT.clade_pearson(side="a")
    assert np.array_equal(again["r"], res["r"], equal_nan=True)


def test_clade_scan_argument_errors():
    SLT, T1, T2 = _slt("gopher_louse")
    with pytest.raises(ValueError):
        SLT.clade_pearson(side="c")
    with pytest.raises(Exception):
        SLT.clade_pearson([T2.size])
    with pytest.raises(Exception):
        SLT.clade_pearson([-1])
    res = SLT.clade_pearson([])
    assert res["r"].shape == (0,) and res["n_links"].shape == (0,)
    # a leaf is a clade of one: its links, no pair unless the leaf has several
    leaf = int(SLT.linklist[0, 0])
    res = SLT.clade_pearson([leaf])
    assert res["n_leafs"][0] == 1 and res["n_links"][0] == int((SLT.linklist[:, 0] == leaf).sum())
