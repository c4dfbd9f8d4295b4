"""CPU: the host side of the per-clade scan (SuchLinkedTrees.clade_pearson / from_linklist):
which link list, intervals and limits reach st_clade_moments, and what is made of the
moments it returns.  The device entry point is replaced by a stand-in that follows its
contract in include/suchtree_b200.h with the oracle's distances; the CUDA path itself is
tests/test_gpu_clades.py."""
import ctypes as C
import os

import numpy as np
import pandas as pd
import pytest
from conftest import DATA, GOLDEN

import oracle as O
from suchtree_b200 import SuchLinkedTrees, SuchTree, _lib, linked, newick


def _host_tree(path):
    """A SuchTree with the host-side state only (no device index)."""
    with open(path) as f:
        ft = newick.flatten(f.read().strip())
    T = SuchTree.__new__(SuchTree)
    T._ft, T._leaves, T._leaf_nodes, T._leaf_ids = ft, ft.leaves, None, None
    T._size, T._n_leaves, T._root, T.device = int(ft.size), int(ft.n_leaves), int(ft.root), 0
    T._handle = None
    T.oracle = O.OracleTree(ft.parent, ft.distance)
    return T


class _StandIn:
    """st_clade_moments as include/suchtree_b200.h specifies it, on the CPU."""

    def __init__(self, real, ta, tb):
        self.real, self.ta, self.tb, self.calls = real, ta, tb, []

    def __getattr__(self, name):
        return getattr(self.real, name)

    # the link-list handle (st_links_create / st_links_destroy): the stand-in keeps the list itself
    def st_links_create(self, ha, hb, ll_p, n_links, h_ref):
        ll = np.ctypeslib.as_array(C.cast(ll_p, C.POINTER(C.c_int64)), (n_links, 2)).copy() if n_links else np.empty((0, 2), np.int64)
        self.handles = getattr(self, "handles", {})
        hid = 1000 + len(self.handles)
        self.handles[hid] = ll
        h_ref._obj.value = hid
        return 0

    def st_links_destroy(self, h):
        self.handles.pop(h.value, None)

    def st_links_clade_moments(self, h, side, lo_p, hi_p, n, min_links, max_links, out_p, nl_p):
        ll = self.handles[h.value]
        return self._clade_moments(ll, side, lo_p, hi_p, n, min_links, max_links, out_p, nl_p)

    def st_clade_moments(self, ha, hb, ll_p, n_links, side, lo_p, hi_p, n, min_links, max_links, out_p, nl_p):
        ll = np.ctypeslib.as_array(C.cast(ll_p, C.POINTER(C.c_int64)), (n_links, 2)) if n_links else np.empty((0, 2), np.int64)
        return self._clade_moments(ll, side, lo_p, hi_p, n, min_links, max_links, out_p, nl_p)

    def _clade_moments(self, ll, side, lo_p, hi_p, n, min_links, max_links, out_p, nl_p):
        lo = np.ctypeslib.as_array(C.cast(lo_p, C.POINTER(C.c_int64)), (n,))
        hi = np.ctypeslib.as_array(C.cast(hi_p, C.POINTER(C.c_int64)), (n,))
        out = C.cast(out_p, C.POINTER(_lib.Moments))
        nl = np.ctypeslib.as_array(C.cast(nl_p, C.POINTER(C.c_int64)), (n,))
        self.calls.append(dict(ll=ll.copy(), side=side, lo=lo.copy(), hi=hi.copy(), min_links=min_links, max_links=max_links))
        min_links = max(min_links, 2)
        for c in range(n):
            sub = ll[(ll[:, side] >= lo[c]) & (ll[:, side] <= hi[c])]
            nl[c] = sub.shape[0]
            for k, _ in _lib.Moments._fields_:
                setattr(out[c], k, 0.0)
            if sub.shape[0] < min_links or (max_links >= 0 and sub.shape[0] > max_links):
                continue
            ids_a, ids_b = O.linked_pairs(sub)
            x, y = self.ta.oracle.distances_f64(ids_a), self.tb.oracle.distances_f64(ids_b)
            m = out[c]
            m.n, m.x0, m.y0 = float(len(x)), float(x[0]), float(y[0])
            x, y = x - x[0], y - y[0]
            m.sx, m.sy, m.sxx, m.syy, m.sxy = x.sum(), y.sum(), (x * x).sum(), (y * y).sum(), (x * y).sum()
        return 0


@pytest.fixture
def slt(monkeypatch):
    def make(name):
        ta, tb, lk = {"gopher_louse": ("test.tree", "lice.tree", "links.csv"),
                      "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"),
                      "arr1": ("arr1_plant.tree", "arr1_animal.tree", "arr1_links.csv")}[name]
        T1, T2 = _host_tree(os.path.join(DATA, ta)), _host_tree(os.path.join(DATA, tb))
        links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
        if set(links.index) != set(T1.leaves.keys()):
            links = links.T
        S = SuchLinkedTrees(T1, T2, links)
        fake = _StandIn(_lib.lib(), T1, T2)
        monkeypatch.setattr(linked._lib, "lib", lambda: fake)
        return S, T1, T2, fake
    return make


@pytest.mark.parametrize("scan", ["b", "a", "ab"])
@pytest.mark.parametrize("name", ["gopher_louse", "fishworm", "arr1"])
def test_clade_pearson_host_side_against_reference_loop(slt, name, scan):
    z = np.load(os.path.join(GOLDEN, "clades.npz"))
    g = {k: z["%s__%s__%s" % (name, scan, k)] for k in ("nodes", "n_leafs", "n_links", "r64")}
    S, T1, T2, fake = slt(name)
    assert np.array_equal(S.linklist, np.load(os.path.join(GOLDEN, "linked_%s.npz" % name))["linklist"])
    if scan == "ab":
        S.subset_a(int(z[name + "__ab__anode"]))
    before = S.linklist.copy()
    res = S.clade_pearson(g["nodes"], side="a" if scan == "a" else "b")
    call = fake.calls[-1]
    assert call["side"] == (1 if scan == "a" else 0) and call["min_links"] == 2 and call["max_links"] == -1
    assert np.array_equal(call["ll"], before)  # the other side's subset in force, reference order
    assert np.array_equal(res["n_leafs"], g["n_leafs"]) and np.array_equal(res["n_links"], g["n_links"])
    ok = np.isfinite(g["r64"])
    assert np.allclose(res["r"][ok], g["r64"][ok], rtol=0, atol=5e-6)
    assert np.all(np.isnan(res["r"][g["n_links"] < 2]))
    assert np.array_equal(S.linklist, before)


def test_scanned_sides_own_subset_is_replaced_per_clade(slt):
    """subset_b(x) then a b scan: each clade's subset_b(node) REPLACES x (as the reference's
    loop does); a subset_a stays in force."""
    S, T1, T2, fake = slt("fishworm")
    full = S.linklist.copy()
    inner = np.nonzero(T2._ft.left != -1)[0]
    S.subset_b(int(inner[3]))
    assert S.subset_n_links < len(full)
    S.clade_pearson(inner[:5], side="b", min_links=4, max_links=77)
    call = fake.calls[-1]
    assert np.array_equal(call["ll"], full) and (call["min_links"], call["max_links"]) == (4, 77)
    res = S.clade_pearson(None, side="a")
    assert np.array_equal(res["node_ids"], np.nonzero(T1._ft.left != -1)[0])
    # the b subset stays in force for an a scan (same links; subset_b lists its columns in
    # get_leaves' breadth-first order, the scan in ascending id: the moments do not care)
    assert sorted(map(tuple, fake.calls[-1]["ll"].tolist())) == sorted(map(tuple, S.linklist.tolist()))


def test_from_linklist_keeps_reference_order_and_validates(slt):
    S, T1, T2, fake = slt("arr1")
    rng = np.random.default_rng(5)
    shuffled = S.linklist[rng.permutation(S.n_links)]
    S2 = SuchLinkedTrees.from_linklist(T1, T2, shuffled)
    assert S2.n_links == S.n_links and np.array_equal(S2.linklist[:, 0], S.linklist[:, 0])
    assert sorted(map(tuple, S2.linklist.tolist())) == sorted(map(tuple, S.linklist.tolist()))
    # stable inside a column: the order given
    first_col = shuffled[shuffled[:, 0] == S2.linklist[0, 0]]
    assert np.array_equal(S2.linklist[: len(first_col)], first_col)
    node = int(np.nonzero(T2._ft.left != -1)[0][2])
    S.subset_b(node)
    S2.subset_b(node)
    assert sorted(map(tuple, S2.linklist.tolist())) == sorted(map(tuple, S.linklist.tolist()))
    assert (S2.subset_b_size, S2.subset_n_links) == (S.subset_b_size, S.subset_n_links)
    with pytest.raises(Exception):
        SuchLinkedTrees.from_linklist(T1, T2, np.array([[1, 0]]))  # 1 is an internal node of TreeB
    with pytest.raises(Exception):
        SuchLinkedTrees.from_linklist(T1, T2, np.array([[0, T1.size]]))
    with pytest.raises(ValueError):
        SuchLinkedTrees.from_linklist(T1, T2, np.zeros((3, 3), np.int64))


def test_clade_intervals_on_deep_and_wide_trees():
    from types import SimpleNamespace

    from suchtree_b200 import synth

    for ft in (synth.caterpillar_tree(3000), synth.balanced_tree(1024), synth.yule_tree(777, seed=3)):
        stub = SimpleNamespace(_size=int(ft.size), _clade_lo_hi=None, _ft=ft)
        lo, hi = SuchTree._clade_intervals(stub)
        depth = O.OracleTree(ft.parent, ft.distance).node_depths()
        for v in np.random.default_rng(0).integers(0, ft.size, 60):
            # the clade of v: the maximal id interval around v of nodes deeper than v
            assert np.all(depth[lo[v]:hi[v] + 1][np.arange(lo[v], hi[v] + 1) != v] > depth[v])
            assert lo[v] == 0 or depth[lo[v] - 1] <= depth[v]
            assert hi[v] == ft.size - 1 or depth[hi[v] + 1] <= depth[v]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_clade_scan_shares_partition_the_clades(slt, world):
    """rank / world: every clade is computed by exactly one rank, with the same result as the
    unsharded scan; the link-pair counts of the shares are close."""
    S, T1, T2, fake = slt("fishworm")
    whole = S.clade_pearson(min_links=4, max_links=150)
    seen, loads = [], []
    for rank in range(world):
        part = S.clade_pearson(min_links=4, max_links=150, rank=rank, world=world)
        idx = np.searchsorted(whole["node_ids"], part["node_ids"])
        assert np.array_equal(whole["node_ids"][idx], part["node_ids"])
        for k in ("n_leafs", "n_links", "n_pairs"):
            assert np.array_equal(whole[k][idx], part[k])
        assert np.array_equal(whole["r"][idx], part["r"], equal_nan=True)
        seen.append(part["node_ids"])
        loads.append(int(part["n_pairs"].sum()))
    assert np.array_equal(np.sort(np.concatenate(seen)), whole["node_ids"])
    assert sum(loads) == int(whole["n_pairs"].sum())
    assert max(loads) - min(loads) <= int(whole["n_pairs"].max())
