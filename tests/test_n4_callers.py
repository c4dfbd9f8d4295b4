"""Thin callers of the device index: distance_to_root, RED,
nearest_neighbors, relationships -- against vectors produced by the unmodified
reference (tests/golden/make_golden_n4.py)."""
import json
import os

import numpy as np
import pytest
from conftest import GOLDEN, tree_source

from suchtree_b200 import SuchTree, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def n4():
    with open(os.path.join(GOLDEN, "n4.json")) as f:
        return json.load(f)


def test_distance_to_root_matches_reference(golden_trees, n4):
    for name, want in n4.items():
        rec = golden_trees[name]
        T = SuchTree(tree_source(name, rec))
        got = np.array([T.distance_to_root(i) for i in range(T.size)])
        ref = np.array(want["distance_to_root"])
        # the reference accumulates in fp32 (MuchTree.pyx:839): depth-scaled tolerance
        l1 = np.abs(got) + 1e-30
        assert np.all(np.abs(got - ref) <= 2e-7 * rec["depth"] * np.maximum(l1, np.abs(ref))), name
        leaf = next(iter(T.leaves))
        assert T.distance_to_root(leaf) == T.distance_to_root(T.leaves[leaf])
        with pytest.warns(DeprecationWarning):
            assert T.get_distance_to_root(leaf) == T.distance_to_root(leaf)


def test_red_matches_reference(golden_trees, n4):
    for name, want in n4.items():
        T = SuchTree(tree_source(name, golden_trees[name]))
        if "red" not in want:
            with pytest.raises(Exception):
                T.relative_evolutionary_divergence
            continue
        red = T.relative_evolutionary_divergence
        assert [int(k) for k in red] == [k for k, _ in want["red"]], name  # preorder insertion order
        got = np.array(list(red.values()))
        ref = np.array([v for _, v in want["red"]])
        # fp32 distances inside the reference; negative branch lengths amplify the difference
        assert np.allclose(got, ref, rtol=2e-5, atol=2e-6), (name, np.abs(got - ref).max())
        assert T.RED is red  # cached, as the reference's SuchTree.RED attribute


def test_nearest_neighbors_match_reference(golden_trees, n4):
    for name, want in n4.items():
        T = SuchTree(tree_source(name, golden_trees[name]))
        for leaf, ref in want["nearest"].items():
            got = T.nearest_neighbors(leaf, k=3)
            assert len(got) == len(ref)
            gd = np.array([d for _, d in got]); rdist = np.array([d for _, d in ref])
            assert np.allclose(gd, rdist, rtol=1e-5, atol=1e-12), (name, leaf)
            # same neighbours wherever the distances are not tied (argsort is not stable across ties)
            for (gn, gdist), (rn, rd_) in zip(got, ref):
                if np.sum(np.isclose(rdist, rd_, rtol=1e-5, atol=1e-12)) == 1:
                    assert gn == rn, (name, leaf)


def test_relationships_is_self_consistent():
    """The reference's relationships() is dead code (its second definition calls a
    missing to_dataframe, MuchTree.pyx:2515-2518); this mirrors the first definition
    (:2158-2179) and is checked against the bulk kernels."""
    T = SuchTree(os.path.join(GOLDEN, "data", "test.tree"))
    df = T.relationships()
    n = T.num_leaves
    assert len(df) == n * (n - 1) // 2
    assert list(df.columns) == ["a", "b", "distance", "a_to_root", "b_to_root", "mrca", "mrca_to_root",
                                "a_to_mrca", "b_to_mrca"]
    for row in df.itertuples():
        assert row.distance == pytest.approx(T.distance(row.a, row.b), rel=1e-12)
        assert row.mrca == T.common_ancestor(row.a, row.b)
        assert row.a_to_mrca + row.b_to_mrca == pytest.approx(row.distance, rel=1e-9, abs=1e-12)
        assert row.mrca_to_root == pytest.approx(T.distance_to_root(int(row.mrca)))


def test_red_on_a_large_tree_is_monotone():
    """Size-independent properties at 100k leaves: RED is 0 at the root, 1 at every
    leaf, and increases from parent to child (positive edge lengths)."""
    ft = synth.yule_tree(100000, seed=1)
    T = SuchTree.from_flat(ft)
    red = T.relative_evolutionary_divergence
    v = np.zeros(T.size)
    v[list(red.keys())] = list(red.values())
    leaves = ft.left == -1
    assert v[T.root_node] == 0 and np.allclose(v[leaves], 1.0)
    nonroot = ft.parent >= 0
    assert np.all(v[nonroot] > v[ft.parent[nonroot]])
