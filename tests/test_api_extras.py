"""The non-accelerated part of the API (suchtree_b200/extras.py) against golden vectors written by
the unmodified reference (tests/golden/make_golden_api.py -> api.json; MuchTree.pyx:370-612,
681-811, 1151-1200, 1424-2251, 3081-3208).

The same checks run twice: on a device-free double (the mixin over the flattened node arrays,
MRCA by a host parent walk) in the CPU suite, and on the real SuchTree / SuchLinkedTrees on the
GPU, where path_between_nodes / distance_to_root go through the device index."""
import json
import os
import warnings

import numpy as np
import pytest

from suchtree_b200 import extras, newick
from suchtree_b200.exceptions import InvalidNodeError, NodeNotFoundError

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "data")
with open(os.path.join(HERE, "golden", "api.json")) as f:
    API = json.load(f)
with open(os.path.join(HERE, "golden", "trees.json")) as f:
    TREES = json.load(f)


def _source(name, text=True):
    """NEWICK text of a golden tree (text=False: what SuchTree() takes -- the string or the path)."""
    nw = TREES[name].get("newick")
    if nw:
        return nw
    path = os.path.join(DATA, name)
    if not text:
        return path
    with open(path) as f:
        return f.read()


def _fl(x):
    return repr(float(x))


def _mat(a):
    a = np.asarray(a)
    return {"shape": list(a.shape), "values": [_fl(v) for v in a.ravel()]}


def _bip(b):
    return sorted(sorted(str(x) for x in side) for side in b)


def _attrs(d):
    return {k: (_fl(v) if isinstance(v, float) else v) for k, v in d.items()}


def _close_mat(got, want, tol=1e-12):
    g = np.asarray(got, dtype=float)
    assert list(g.shape) == want["shape"]
    w = np.array([float(v) for v in want["values"]]).reshape(want["shape"])
    assert np.allclose(g, w, rtol=tol, atol=tol)


def check_tree(T, r):
    n = T.size
    assert n == r["size"]
    assert [list(T.get_ancestors(i)) for i in range(n)] == r["ancestors"]
    assert [_fl(T.get_support(i)) for i in range(n)] == r["support"]
    sub = r["sub"]
    assert [int(i) for i in T.get_nodes()] == r["nodes"]
    assert [int(i) for i in T.get_internal_nodes()] == r["internal"]
    assert [int(i) for i in T.get_nodes(sub)] == r["nodes_sub"]
    assert [int(i) for i in T.get_internal_nodes(sub)] == r["internal_sub"]
    assert isinstance(T.get_nodes(), np.ndarray)
    pairs = [tuple(p) for p in r["pairs"]]
    assert [bool(T.is_descendant(a, b)) for a, b in pairs] == r["is_descendant"]
    assert [int(T.is_ancestor(a, b)) for a, b in pairs] == r["is_ancestor"]
    assert [bool(T.is_sibling(a, b)) for a, b in pairs] == r["is_sibling"]
    assert [bool(T.is_sibling(*T.get_children(i))) for i in r["internal"]] == r["sibling_of_children"]
    assert [bool(T.has_children(i)) for i in range(n)] == r["has_children"]
    assert [bool(T.has_parent(i)) for i in range(n)] == r["has_parent"]
    assert [[int(x) for x in T.path_between_nodes(a, b)] for a, b in pairs] == r["paths"]
    assert [_bip(b) for b in T.bipartitions(by_id=True)] == r["bipartitions_by_id"]
    assert [_bip(b) for b in T.bipartitions()] == r["bipartitions"]
    b0 = T.bipartition(r["internal"][0])
    assert isinstance(b0, frozenset) and all(isinstance(s, frozenset) for s in b0)
    assert [[int(i), _fl(d)] for i, d in T.traverse_inorder()] == r["inorder"]
    assert [int(i) for i in T.traverse_inorder(include_distances=False)] == r["inorder_ids"]
    for name in ("preorder", "postorder", "levelorder", "leaves_only", "internal_only"):
        fn = getattr(T, "traverse_" + name)
        assert [int(i) for i in fn()] == r[name], name
        assert [int(i) for i in fn(sub)] == r[name + "_sub"], name
    assert [[int(i), int(d)] for i, d in T.traverse_with_depth()] == r["with_depth"]
    assert [[int(i), int(d)] for i, d in T.traverse_with_depth(sub)] == r["with_depth_sub"]
    assert [[int(i), _fl(d), _fl(c)] for i, d, c in T.traverse_with_distances()] == r["with_distances"]
    assert [[int(i), _fl(d), _fl(c)] for i, d, c in T.traverse_with_distances(sub)] == r["with_distances_sub"]
    for tag, start in (("", None), ("_sub", sub)):
        A = T.adjacency_matrix(start)
        assert [int(i) for i in A["node_ids"]] == r["adjacency" + tag]["node_ids"]
        assert _mat(A["adjacency_matrix"]) == r["adjacency" + tag]["matrix"]
        Lp = T.laplacian_matrix(start)
        assert [int(i) for i in Lp["node_ids"]] == r["laplacian" + tag]["node_ids"]
        _close_mat(Lp["laplacian"], r["laplacian" + tag]["matrix"])
        D = T.degree_sequence(start)
        assert [int(i) for i in D["degrees"]] == r["degrees" + tag]["degrees"]
        assert int(D["max_degree"]) == r["degrees" + tag]["max"] and int(D["min_degree"]) == r["degrees" + tag]["min"]
    inc = T.incidence_matrix()
    assert list(inc["incidence_matrix"].shape) == r["incidence"]["shape"]
    assert [int(v) for v in inc["incidence_matrix"].ravel()] == r["incidence"]["matrix"]
    assert [int(i) for i in inc["node_ids"]] == r["incidence"]["node_ids"]
    assert [[int(a), int(b)] for a, b in inc["edge_list"]] == r["incidence"]["edge_list"]
    if r["incidence_sub_error"]:
        with pytest.raises(IndexError):
            T.incidence_matrix(sub)
    assert [[int(c), int(p), _attrs(a)] for c, p, a in T.to_networkx_edges()] == r["nx_edges"]

    def nodes_no_rd(recs):  # distance_to_root is compared separately (fp64 sum here, fp32 upstream)
        return [[i, {k: v for k, v in a.items() if k != "distance_to_root"}] for i, a in recs]

    got = [[int(i), _attrs(a)] for i, a in T.to_networkx_nodes()]
    assert nodes_no_rd(got) == nodes_no_rd(r["nx_nodes"])
    for (_, a), (_, b) in zip(got, r["nx_nodes"]):
        assert abs(float(a["distance_to_root"]) - float(b["distance_to_root"])) <= 1e-5 * max(1.0, abs(float(b["distance_to_root"])))
    assert nodes_no_rd([[int(i), _attrs(a)] for i, a in T.to_networkx_nodes(sub)]) == nodes_no_rd(r["nx_nodes_sub"])
    assert T.to_newick() == r["newick"]
    assert T.to_newick(include_support=False, include_distances=False) == r["newick_plain"]
    assert T.to_newick(sub) == r["newick_sub"]
    leaf = int(T.leaf_node_ids[0])
    for what, fn in (("bipartition_of_leaf", lambda: T.bipartition(leaf)),
                     ("ancestors_out_of_range", lambda: list(T.get_ancestors(n))),
                     ("support_unknown_name", lambda: T.get_support("no such leaf")),
                     ("path_bad_type", lambda: T.path_between_nodes(1.5, 0))):
        want = r["errors"][what]
        with pytest.raises(Exception) as ei:
            fn()
        assert [type(ei.value).__name__, str(ei.value)] == want, what
    # deprecated aliases keep their warnings and results
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert list(T.get_lineage(leaf)) == r["ancestors"][leaf]
        assert [int(i) for i in T.in_order(distances=False)] == r["inorder_ids"]
        assert _bip(T.get_bipartition(r["internal"][0], by_id=True)) == r["bipartitions_by_id"][0]
        assert T.is_internal_node(r["internal"][0]) is True
        assert [int(i) for i in T.get_leafs(sub)] == [int(i) for i in T.get_leaves(sub)]
        assert _mat(T.adjacency()["adjacency_matrix"]) == r["adjacency"]["matrix"]
        assert [int(c) for c, _, _ in T.edges_data()] == [e[0] for e in r["nx_edges"]]
    assert sum(issubclass(x.category, DeprecationWarning) for x in w) >= 7
    # leaf links live in a side table
    T.link_leaf(leaf, 5)
    assert T.get_links([leaf]).tolist() == [5]
    with pytest.raises(Exception):
        T.link_leaf(r["internal"][0], 1)


# ------------------------------------------------------------------ CPU: device-free double --
class HostTree(extras.TreeExtras):
    """The mixin over the flattened arrays alone; the few methods it borrows from SuchTree are
    restated with host walks."""

    def __init__(self, text):
        from suchtree_b200.tree import SuchTree

        ft = newick.flatten_py(text)
        self._ft, self._root, self._size = ft, int(ft.root), int(ft.size)
        self.leaves = ft.leaves
        self.leaf_nodes = {v: k for k, v in ft.leaves.items()}
        self.size = self._size
        self.polytomy_epsilon = float(np.finfo(np.float64).eps)
        for name in ("_validate_node", "_validate_node_pair", "get_children", "get_leaves", "get_descendants",
                     "traverse_preorder", "is_leaf", "is_internal", "is_root", "_clade_interval", "is_ancestor"):
            setattr(self, name, getattr(SuchTree, name).__get__(self))

    @property
    def leaf_node_ids(self):
        return np.array(list(self.leaves.values()))

    def common_ancestor(self, a, b):
        anc = {a} | set(self.get_ancestors(a))
        while b not in anc:
            b = int(self._ft.parent[b])
        return b

    def distance_to_root(self, node):
        d, i = 0.0, node
        while self._ft.parent[i] != -1:
            d += float(self._ft.distance[i])
            i = int(self._ft.parent[i])
        return d


@pytest.mark.parametrize("name", sorted(API["trees"]))
def test_extras_on_host_double(name):
    check_tree(HostTree(_source(name)), API["trees"][name])


# ------------------------------------------------------------------ GPU: the real classes -----
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(API["trees"]))
def test_extras_on_device_tree(name):
    from suchtree_b200 import SuchTree

    check_tree(SuchTree(_source(name, text=False)), API["trees"][name])


@pytest.mark.gpu
@pytest.mark.parametrize("name,files", [("gopher_louse", ("test.tree", "lice.tree", "links.csv")),
                                        ("fishworm", ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"))])
def test_linked_graph_matrices_and_spectrum(name, files):
    import pandas as pd

    from suchtree_b200 import SuchLinkedTrees, SuchTree

    r = API["linked"][name]
    T1, T2 = SuchTree(os.path.join(DATA, files[0])), SuchTree(os.path.join(DATA, files[1]))
    links = pd.read_csv(os.path.join(DATA, files[2]), index_col=0)
    if set(links.index) != set(T1.leaves.keys()):
        links = links.T
    SLT = SuchLinkedTrees(T1, T2, links)

    def spectrum_close(got, want):
        w = np.array([float(v) for v in want])
        assert got.shape == w.shape and np.allclose(got, w, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(w).max())))

    aj = SLT.adjacency()
    if "adjacency" in r:
        _close_mat(aj, r["adjacency"], tol=1e-13)
        _close_mat(SLT.laplacian(), r["laplacian"], tol=1e-13)
    else:
        assert list(aj.shape) == r["adjacency_shape"] and int(np.count_nonzero(aj)) == r["adjacency_nonzeros"]
        assert np.allclose(aj.sum(axis=1), [float(v) for v in r["adjacency_row_sums"]], rtol=1e-12, atol=1e-15)
    assert np.array_equal(aj, aj.T)
    spectrum_close(SLT.spectrum(), r["spectrum"])
    SLT.subset_a(r["subset"]["a"])
    SLT.subset_b(r["subset"]["b"])
    assert SLT.subset_n_links == r["subset"]["n_links"]
    if "adjacency_shape" in r["subset"]:
        aj = SLT.adjacency()
        assert list(aj.shape) == r["subset"]["adjacency_shape"]
        assert np.allclose(aj.sum(axis=1), [float(v) for v in r["subset"]["adjacency_row_sums"]], rtol=1e-12, atol=1e-15)
        spectrum_close(SLT.spectrum(), r["subset"]["spectrum"])


@pytest.mark.gpu
def test_newick_round_trip_and_deep_ladder():
    """to_newick() of a tree re-read gives the same arrays; a 200,000-leaf ladder neither
    recurses nor takes long."""
    from suchtree_b200 import SuchTree, synth

    T = SuchTree(os.path.join(DATA, "test.tree"))
    U = SuchTree(T.to_newick())
    assert U.size == T.size and U.leaves == T.leaves
    assert np.array_equal(U._ft.parent, T._ft.parent) and np.array_equal(U._ft.distance, T._ft.distance)
    L = SuchTree.from_flat(synth.caterpillar_tree(200000, seed=3))
    text = L.to_newick(include_distances=False)
    assert text.count("(") == 199999 and text.endswith(";")
    assert sum(1 for _ in L.traverse_postorder()) == L.size
    assert sum(1 for _ in L.get_ancestors(0)) == L.depth - 1


def test_exceptions_carry_reference_messages():
    e = InvalidNodeError(3, message="Node 3 is not a leaf node")
    assert str(e) == "Node 3 is not a leaf node" and e.node_id == 3
    assert str(NodeNotFoundError("x")) == "Leaf name not found: x."
