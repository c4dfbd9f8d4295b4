"""GPU parity for the all-pairs matrix writer, the linked-tree entry points, the
samplers and pearson()."""
import os

import numpy as np
import pandas as pd
import pytest
from conftest import DATA, GOLDEN, tree_source

import oracle as O
from suchtree_b200 import philox_host as philox_ref
from suchtree_b200 import SuchLinkedTrees, SuchTree, moments_pearson, pearson, synth

pytestmark = pytest.mark.gpu


def _pairs_of(ids):
    ids = np.asarray(ids, dtype=np.int64)
    a, b = np.meshgrid(ids, ids, indexing="ij")
    return np.stack([a.ravel(), b.ravel()], axis=1)


# ------------------------------------------------------------------ matrix ---
@pytest.mark.parametrize("geom", [(0, 0), (2, 1), (4, 2)])
def test_matrix_small_trees(golden_trees, geom):
    for name, rec in golden_trees.items():
        T = SuchTree(tree_source(name, rec), _block_shift=geom[0], _micro_shift=geom[1])
        D = T.pairwise_distances()
        ids = T.leaf_node_ids
        n = len(ids)
        assert D.shape == (n, n)
        want = T.distances_bulk(_pairs_of(ids)).reshape(n, n)
        assert np.array_equal(D, want), name
        assert np.array_equal(D, D.T) and not D.diagonal().any(), name
        # arbitrary node list: internal nodes, unsorted, a duplicate
        rng = np.random.default_rng(3)
        nodes = [int(x) for x in rng.integers(0, T.size, size=min(T.size, 11))] + [T.root_node]
        nodes.append(nodes[0])
        D2 = T.pairwise_distances(nodes)
        assert np.array_equal(D2, T.distances_bulk(_pairs_of(nodes)).reshape(len(nodes), len(nodes))), name


def test_matrix_reference_api():
    """test_new_api.py:410-423, 708-727."""
    T = SuchTree(os.path.join(DATA, "test.tree"))
    leaves = list(T.leaves.keys())[:5]
    M = T.pairwise_distances(leaves)
    assert M.shape == (5, 5) and np.allclose(M, M.T) and np.allclose(np.diag(M), 0)
    assert M[0, 1] == pytest.approx(T.distance(leaves[0], leaves[1]), abs=1e-12)
    r = T.distance_matrix(leaves)
    assert set(r) == {"distance_matrix", "node_ids", "node_names"} and r["node_names"] == leaves
    r = T.distance_matrix()
    assert r["distance_matrix"].shape == (15, 15) and list(r["node_ids"]) == list(T.leaves.values())
    r = T.distance_matrix([1, 0])
    assert r["node_names"] == ["node_1", T.leaf_nodes[0]]
    nn = T.nearest_neighbors("Ttal", k=3)
    assert len(nn) == 3 and nn[0][1] <= nn[1][1] <= nn[2][1]


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("gen,n_leaves", [(synth.yule_tree, 3001), (synth.caterpillar_tree, 2500), (synth.balanced_tree, 4096)])
def test_matrix_mid_size_exact(gen, n_leaves, wide):
    """Every tile class (ordered upper/lower, diagonal) against the pair kernel; odd n
    exercises the unaligned store path.  Synthetic edges: exact, so bitwise equality
    and exact symmetry."""
    ft = gen(n_leaves, seed=4)
    T = SuchTree.from_flat(ft, _wide=wide)
    D = T.pairwise_distances()
    ids = np.arange(0, 2 * n_leaves, 2, dtype=np.int64)
    want = T.distances_bulk(_pairs_of(ids)).reshape(n_leaves, n_leaves)
    assert np.array_equal(D, want)
    assert np.array_equal(D, D.T)
    ot = O.OracleTree(ft.parent, ft.distance)
    rng = np.random.default_rng(5)
    ij = rng.integers(0, n_leaves, size=(3000, 2))
    assert np.array_equal(D[ij[:, 0], ij[:, 1]], ot.distances_f64_climb(2 * ij))
    # sorted list of arbitrary nodes takes the ordered-tile path too
    nodes = np.unique(rng.integers(0, ft.size, size=1500))
    D3 = T.pairwise_distances([int(x) for x in nodes])
    assert np.array_equal(D3, T.distances_bulk(_pairs_of(nodes)).reshape(len(nodes), len(nodes)))


def test_matrix_real_tree_subset_tolerance():
    """Real edge lengths (ml.tree, epsilon edges included): matrix vs O2 to 1e-12."""
    T = SuchTree(os.path.join(DATA, "ml.tree.gz"))
    ids = T.leaf_node_ids[:1500].astype(np.int64)
    D = T.pairwise_distances([int(x) for x in ids])
    ot = O.OracleTree(T._ft.parent, T._ft.distance, depth=T.depth)
    P = _pairs_of(ids)
    sel = np.random.default_rng(6).integers(0, P.shape[0], size=20000)
    want, l1 = ot.distances_f64(P[sel], with_l1=True)
    assert np.all(np.abs(D.ravel()[sel] - want) <= 1e-12 * l1 + 1e-300)
    assert np.allclose(D, D.T, rtol=1e-15, atol=0)


def test_matrix_row_sharding_on_device():
    """Row blocks written by separate calls (the multi-GPU decomposition) tile the
    same matrix; device-resident output."""
    import ctypes as C

    import torch

    from suchtree_b200 import _lib

    ft = synth.yule_tree(2000, seed=8)
    T = SuchTree.from_flat(ft)
    n = 2000
    full = T.pairwise_distances()
    dev = torch.device("cuda", T.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    parts = []
    for r0, r1 in ((0, 500), (500, 1337), (1337, 2000)):
        out = torch.empty((r1 - r0, n), dtype=torch.float64, device=dev)
        rc = _lib.lib().st_distance_matrix(T._handle, None, n, r0, r1, out.data_ptr(), 1, stream)
        _lib.check(rc)
        torch.cuda.synchronize()
        parts.append(out.cpu().numpy())
    assert np.array_equal(np.concatenate(parts), full)


# ------------------------------------------------------------------ linked ---
LINKED = {
    "gopher_louse": ("test.tree", "lice.tree", "links.csv"),
    "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"),
    "perfect0": ("perfect0_host.tree", "perfect0_guest.tree", "perfect0_links.csv"),
    "arr1": ("arr1_plant.tree", "arr1_animal.tree", "arr1_links.csv"),
}


def _slt(name, seed=12345):
    ta, tb, lk = LINKED[name]
    T1 = SuchTree(os.path.join(DATA, ta))
    T2 = SuchTree(os.path.join(DATA, tb))
    links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
    if set(links.index) != set(T1.leaves.keys()):
        links = links.T
    np.random.seed(seed)
    return SuchLinkedTrees(T1, T2, links), T1, T2


@pytest.mark.parametrize("name", list(LINKED))
def test_linked_distances_match_reference(name):
    z = np.load(os.path.join(GOLDEN, "linked_%s.npz" % name))
    SLT, T1, T2 = _slt(name)
    assert np.array_equal(SLT.linklist, z["linklist"]) and SLT.n_links == int(z["n_links"])
    r = SLT.linked_distances()
    assert np.array_equal(r["ids_A"], z["ids_A"]) and np.array_equal(r["ids_B"], z["ids_B"])
    assert r["n_pairs"] == int(z["n_pairs"]) == r["n_samples"] and r["deviation_a"] is None
    for key, T, ids in (("TreeA", T1, z["ids_A"]), ("TreeB", T2, z["ids_B"])):
        ot = O.OracleTree(T._ft.parent, T._ft.distance, depth=T.depth)
        want, l1 = ot.distances_f64(ids, with_l1=True)
        assert np.all(np.abs(r[key] - want) <= 1e-12 * l1 + 1e-300)
        assert np.all(np.abs(r[key] - z[key]) <= 2e-7 * T.depth * l1 + 1e-300)  # raw reference (fp32 sums)
    # pearson(): fp64 on the device vs the reference's fp32-accumulated value
    assert pearson(r["TreeA"], r["TreeB"]) == pytest.approx(O.pearson_f64(r["TreeA"], r["TreeB"]), abs=1e-12)
    assert pearson(r["TreeA"], r["TreeB"]) == pytest.approx(float(z["pearson"]), abs=2e-4)


@pytest.mark.parametrize("name", list(LINKED))
def test_linked_pearson_fused_equals_two_step(name):
    """linked_pearson() (moments fused on the device, nothing materialised) against
    pearson(linked_distances()) and the reference's own value; sharded pair ranges
    (the multi-GPU decomposition) add up to the same moments."""
    z = np.load(os.path.join(GOLDEN, "linked_%s.npz" % name))
    SLT, T1, T2 = _slt(name)
    r = SLT.linked_distances()
    want = O.pearson_f64(r["TreeA"], r["TreeB"])
    assert SLT.linked_pearson() == pytest.approx(want, abs=1e-10)
    assert SLT.linked_pearson() == pytest.approx(float(z["pearson"]), abs=2e-4)
    total = r["n_pairs"]
    cut = [0, total // 3, total // 3 + 1, total]
    parts = [SLT.linked_moments(first_pair=a, n_pairs=b - a) for a, b in zip(cut[:-1], cut[1:])]
    whole = SLT.linked_moments()
    for k in ("n", "sx", "sy", "sxx", "syy", "sxy"):
        assert sum(getattr(p, k) for p in parts) == pytest.approx(getattr(whole, k), rel=1e-12, abs=1e-12)
    assert whole.sx == pytest.approx(r["TreeA"].sum(), rel=1e-12) and whole.sxy == pytest.approx(
        (r["TreeA"] * r["TreeB"]).sum(), rel=1e-12)
    # clade scan step: subset, then the fused correlation
    node = int(T2.get_parent(list(T2.leaves.values())[0]))
    if T2.get_parent(node) >= 0:
        node = int(T2.get_parent(node))
    SLT.subset_b(node)
    if SLT.subset_n_links > 2:
        rs = SLT.linked_distances()
        assert SLT.linked_pearson() == pytest.approx(O.pearson_f64(rs["TreeA"], rs["TreeB"]), abs=1e-9)


def test_published_gopher_louse_pearson(published):
    SLT, _, _ = _slt("gopher_louse")
    r = SLT.linked_distances()
    assert pearson(r["TreeA"], r["TreeB"]) == pytest.approx(published["gopher_louse_pearsonr"]["value"], abs=2e-7)
    assert {tuple(x) for x in SLT.linklist.tolist()} == {tuple(x) for x in published["gopher_louse_linklist"]["value"]}


@pytest.mark.parametrize("name", ["gopher_louse", "fishworm", "arr1"])
def test_sample_linked_distances_reproduces_reference_stream(name):
    """Same numpy seed -> same xorshift64* seed -> the device draws the reference's
    link pairs: identical sample count, distances equal up to the reference's fp32
    accumulation, same convergence decision."""
    s = np.load(os.path.join(GOLDEN, "sampled_%s.npz" % name))
    SLT, T1, T2 = _slt(name)
    assert SLT._seed == int(s["rng_seed"])
    r = SLT.sample_linked_distances(sigma=float(s["sigma"]), buckets=int(s["buckets"]), n=int(s["n"]),
                                    maxcycles=int(s["maxcycles"]))
    assert r is not None and r["n_samples"] == int(s["n_samples"]) and r["n_pairs"] == float(s["n_pairs"])
    assert r["TreeA"].shape == s["TreeA"].shape
    tol = 2e-7 * max(T1.depth, T2.depth)
    assert np.allclose(r["TreeA"], s["TreeA"], rtol=tol, atol=1e-12)
    assert np.allclose(r["TreeB"], s["TreeB"], rtol=tol, atol=1e-12)
    assert r["deviation_a"] == pytest.approx(float(s["deviation_a"]), rel=1e-3, abs=1e-6)
    assert r["deviation_b"] == pytest.approx(float(s["deviation_b"]), rel=1e-3, abs=1e-6)
    # and exactly against the oracle's restatement of the stream + O2 distances
    seed = int(s["rng_seed"])
    ll = np.ascontiguousarray(SLT.linklist)
    otA = O.OracleTree(T1._ft.parent, T1._ft.distance, depth=T1.depth)
    otB = O.OracleTree(T2._ft.parent, T2._ft.distance, depth=T2.depth)
    qa_all, qb_all = [], []
    for _ in range(r["n_samples"] // int(s["n"])):
        qa, qb, seed = O.sample_bucket(seed, ll, int(s["n"]))
        qa_all.append(qa)
        qb_all.append(qb)
    wa, l1a = otA.distances_f64(np.concatenate(qa_all), with_l1=True)
    wb, l1b = otB.distances_f64(np.concatenate(qb_all), with_l1=True)
    assert np.all(np.abs(r["TreeA"] - wa) <= 1e-12 * l1a + 1e-300)
    assert np.all(np.abs(r["TreeB"] - wb) <= 1e-12 * l1b + 1e-300)
    assert SLT._seed == seed  # the state advanced exactly as the reference's would


def test_link_bookkeeping_accessors():
    """Per-column access, the dense matrix and the link list against the DataFrame they
    were built from (what SuchTree/tests/test_SuchLinkedTrees.py:129-246 checks), with
    shuffled rows so that matrix order and tree order differ."""
    T = SuchTree(os.path.join(DATA, "test.tree"))
    N = T.num_leaves
    rng = np.random.default_rng(17)
    row_names = list(T.leaves.keys())
    rng.shuffle(row_names)
    links = pd.DataFrame(rng.integers(0, 3, size=(N, N)), columns=list(T.leaves.keys()), index=row_names)
    SLT = SuchLinkedTrees(T, T, links)
    assert SLT.n_rows == SLT.n_cols == N and SLT.n_links == int((links.to_numpy() > 0).sum())
    assert list(SLT.col_ids) == list(T.leaves.values()) and SLT.row_names == list(T.leaves.keys())
    mask = links > 0
    for n, colname in enumerate(links.columns):
        s = mask[colname]
        want = {T.leaves[x] for x in s[s].index}
        assert set(SLT.get_column_leafs(n)) == want == set(SLT.get_column_leafs(colname))
        rows = {list(SLT.row_ids).index(v) for v in want}
        assert set(SLT.get_column_leafs(n, as_row_ids=True)) == rows
        c = SLT.get_column_links(n)
        assert [bool(s[r]) for r in SLT.row_names] == c.tolist()
    lm = SLT.linkmatrix
    for ci, col in enumerate(SLT.col_names):
        for ri, row in enumerate(SLT.row_names):
            assert bool(links.at[row, col]) == lm[ri][ci]
    un = links.unstack()
    want = {(T.leaves[c], T.leaves[r]) for c, r in un[un > 0].index}
    assert {(int(b), int(a)) for b, a in SLT.linklist} == want
    with pytest.raises(Exception):
        SLT.get_column_leafs(N + 1)


def test_sampler_returns_none_at_maxcycles():
    SLT, _, _ = _slt("gopher_louse")
    assert SLT.sample_linked_distances(sigma=1e-9, buckets=4, n=16, maxcycles=2) is None


def test_subsets_follow_reference_order():
    SLT, T1, T2 = _slt("fishworm")
    full = SLT.linklist.copy()
    node = T2.get_parent(int(full[0, 0]))
    if T2.get_parent(node) != -1:
        node = T2.get_parent(node)
    SLT.subset_b(node)
    leaves_b = set(T2.get_leaves(node).tolist())
    assert set(SLT.linklist[:, 0].tolist()) <= leaves_b and SLT.subset_n_links == sum(x in leaves_b for x in full[:, 0])
    SLT.subset_b(T2.root_node)
    assert SLT.subset_n_links == len(full)
    node_a = T1.get_parent(int(full[0, 1]))
    SLT.subset_a(node_a)
    leaves_a = set(T1.get_leaves(node_a).tolist())
    assert set(SLT.linklist[:, 1].tolist()) <= leaves_a
    r = SLT.linked_distances()
    assert r["n_pairs"] == SLT.subset_n_links * (SLT.subset_n_links - 1) // 2


# ------------------------------------------------------ Philox moments path --
def test_sample_moments_against_restated_stream():
    """Known-answer check of the fused sampler: identical Philox index stream restated
    in numpy + oracle distances; |dr| <= 1e-9 (SURVEY §8d cfg 4)."""
    fa, fb = synth.yule_tree(4000, seed=4), synth.yule_tree(4000, seed=5)
    TA, TB = SuchTree.from_flat(fa), SuchTree.from_flat(fb)
    rng = np.random.default_rng(6)
    L = 5000
    ll = np.stack([2 * rng.integers(0, 4000, L), 2 * rng.integers(0, 4000, L)], axis=1).astype(np.int64)
    names_a = list(TA.leaves.keys())
    SLT = SuchLinkedTrees.__new__(SuchLinkedTrees)
    SLT._TreeA, SLT._TreeB = TA, TB
    SLT._np_linklist, SLT._subset_n_links = ll, L
    n = 200001
    m = SLT.sample_moments(n, seed=7, first_sample=0, x0=30.0, y0=30.0)
    l1, l2 = philox_ref.sampled_links(L, 7, 0, n)
    oa, ob = O.OracleTree(fa.parent, fa.distance), O.OracleTree(fb.parent, fb.distance)
    x = oa.distances_f64_climb(np.stack([ll[l1, 1], ll[l2, 1]], axis=1))
    y = ob.distances_f64_climb(np.stack([ll[l1, 0], ll[l2, 0]], axis=1))
    assert m.n == n
    assert m.sx == pytest.approx(float((x - 30.0).sum()), rel=1e-12)
    assert m.sxy == pytest.approx(float(((x - 30.0) * (y - 30.0)).sum()), rel=1e-11)
    assert moments_pearson(m) == pytest.approx(O.pearson_f64(x, y), abs=1e-9)
    # sharded over "ranks": disjoint sample ranges, sums add up (the NCCL all-reduce payload)
    parts = [SLT.sample_moments(c, seed=7, first_sample=f, x0=30.0, y0=30.0)
             for f, c in ((0, 70001), (70001, 29999), (100000, 100001))]
    tot = type(m)()
    tot.n, tot.x0, tot.y0 = n, 30.0, 30.0
    for k in ("sx", "sy", "sxx", "syy", "sxy"):
        setattr(tot, k, sum(getattr(p, k) for p in parts))
    assert moments_pearson(tot) == pytest.approx(moments_pearson(m), abs=1e-12)
    # identical trees + bijective links -> r = 1
    SLT2 = SuchLinkedTrees.__new__(SuchLinkedTrees)
    SLT2._TreeA, SLT2._TreeB = TA, TA
    ids = 2 * np.arange(4000, dtype=np.int64)
    SLT2._np_linklist, SLT2._subset_n_links = np.stack([ids, ids], axis=1), 4000
    assert SLT2.sample_pearson(100000, seed=1) == pytest.approx(1.0, abs=1e-12)


def test_pearson_api_and_vectors():
    import json

    with open(os.path.join(GOLDEN, "pearson_vectors.json")) as f:
        pv = json.load(f)
    rng = np.random.default_rng(7)
    for n, rhex in zip(pv["n"], pv["r"]):
        x = rng.random(n)
        y = 0.3 * x + rng.random(n)
        got = pearson(x, y)
        assert got == pytest.approx(O.pearson_f64(x, y), abs=1e-12)
        assert got == pytest.approx(float.fromhex(rhex), abs=2e-3 if n < 4 else 5e-5)  # reference: fp32 sums
    with pytest.raises(Exception):
        pearson(np.zeros(3), np.zeros(4))
    with pytest.raises(ValueError):
        pearson(np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32))
    assert pearson(np.zeros(0), np.zeros(0)) == 0.0
    big = np.random.default_rng(1).random(3_000_001)
    assert pearson(big, 2 * big + 1) == pytest.approx(1.0, abs=1e-12)
