"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on
the same seeded inputs, against the golden vectors produced by the unmodified
reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.

Tolerances (north_star): MRCA ids bit-exact; distances within 1e-12 RELATIVE of
the reference's path summation in fp64 over its fp32-quantised edges (oracle O2),
measured against the path's L1 norm so that trees with negative branch lengths
(nj.tree) are covered; against the raw reference output (fp32 accumulator,
MuchTree.pyx:924) a depth-scaled fp32 tolerance."""
import os

import numpy as np
import pytest
from conftest import GOLDEN, load_pairs, matrix_rows, read_tree_text, tree_source

import oracle as O
from suchtree_b200 import philox_host as philox_ref
from suchtree_b200 import SuchTree, synth

pytestmark = pytest.mark.gpu

REL = 1e-12
# (block_shift, micro_shift): defaults, and tiny blocks so that small trees reach the
# block-table fast path, the micro-table path and the in-micro-block scan
GEOMETRIES = [(0, 0), (2, 1), (3, 2), (4, 2), (5, 3)]


def _check_distances(got, want64, l1, what):
    err = np.abs(got - want64)
    bad = err > REL * l1 + 1e-300
    assert not bad.any(), "%s: %d distances beyond 1e-12 (worst %.3e rel)" % (
        what, bad.sum(), (err / np.maximum(l1, 1e-300)).max())


def _all_pairs(n):
    return np.array([(a, b) for a in range(n) for b in range(n)], dtype=np.int64)


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("geom", GEOMETRIES)
def test_small_trees_against_reference_and_oracle(golden_trees, geom, wide):
    layouts = set()
    for name, rec in golden_trees.items():
        T = SuchTree(tree_source(name, rec), _block_shift=geom[0], _micro_shift=geom[1], _wide=wide)
        layouts.add(T.index_info["layout"])
        assert (T.size, T.depth, T.num_leaves, T.root_node) == (
            rec["size"], rec["depth"], rec["num_leaves"], rec["root"]), name
        assert T.leaves == rec["leaves"], name
        z = load_pairs(name)
        ot = O.OracleTree.from_newick(tree_source(name, rec))
        # device-built per-node arrays
        depth, hi, lo = T.export_index()
        assert np.array_equal(depth, ot.node_depths()), name
        # MRCA: bit-exact against the reference itself
        assert np.array_equal(T.common_ancestors_bulk(z["pairs"]), z["mrca"]), name
        # distances: 1e-12 against O2, fp32-scaled against the raw reference output
        d64, l1 = ot.distances_f64(z["pairs"], with_l1=True)
        got = T.distances_bulk(z["pairs"])
        _check_distances(got, d64, l1, name)
        assert np.all(np.abs(got - z["distance"]) <= 2e-7 * rec["depth"] * l1 + 1e-300), name
        # root distance itself (pairs (v, root))
        pr = np.stack([np.arange(T.size), np.full(T.size, T.root_node)], axis=1).astype(np.int64)
        rd64, rl1 = ot.distances_f64(pr, with_l1=True)
        _check_distances(hi + lo, rd64, rl1, name + " rd")
    # the fixtures exercise every layout: epsilon edges keep the wide records (layout 2:
    # wide records + 32-bit block tables), exact trees get the compact ones (layout 1)
    assert layouts == ({0} if wide else {1, 2})


@pytest.mark.parametrize("name", ["ml", "nj"])
def test_big_real_trees(name, bigtrees):
    z = np.load(os.path.join(GOLDEN, "pairs_big_%s.npz" % name))
    T = SuchTree(os.path.join(GOLDEN, "data", "%s.tree.gz" % name))
    info = bigtrees[name]
    assert (T.size, T.depth, T.num_leaves, T.root_node) == (
        info["size"], info["depth"], info["num_leaves"], info["root"])
    assert np.array_equal(T.common_ancestors_bulk(z["pairs"]), z["mrca"])
    ot = O.OracleTree(T._ft.parent, T._ft.distance, depth=T.depth)
    d64, l1 = ot.distances_f64(z["pairs"], with_l1=True)
    got = T.distances_bulk(z["pairs"])
    _check_distances(got, d64, l1, name)
    assert np.all(np.abs(got - z["distance"]) <= 2e-7 * info["depth"] * l1 + 1e-300)


def test_reference_test_matrix():
    """SuchTree/tests/test_SuchTree.py:56-88 and test_new_api.py:362-408."""
    T = SuchTree(os.path.join(GOLDEN, "data", "test.tree"))
    rows = matrix_rows()
    for a, b, d1 in rows:
        assert d1 == pytest.approx(T.distance(a, b), rel=1e-3, abs=1e-3)
    ids = np.array([(T.leaves[a], T.leaves[b]) for a, b, _ in rows], dtype=np.int64)
    want = [d for _, _, d in rows]
    with pytest.warns(DeprecationWarning):
        got = T.distances(ids)
    assert got == pytest.approx(want, rel=1e-3, abs=1e-3)
    assert T.distances_by_name([(a, b) for a, b, _ in rows]) == pytest.approx(want, rel=1e-3, abs=1e-3)


@pytest.mark.parametrize(
    "gen,n_leaves,n_pairs,literal",
    [
        (synth.yule_tree, 100000, 1000000, True),      # cfg 2 tree, 1e6 pairs: seconds for the oracle
        (synth.balanced_tree, 1 << 17, 300000, True),
        (synth.balanced_tree, 100003, 100000, True),
        (synth.caterpillar_tree, 3000, 5000, True),     # O(depth^2) literal oracle still feasible
        (synth.caterpillar_tree, 1000000, 600, False),   # cfg 3 deep-path worst case: climbing oracle
        (synth.yule_tree, 1000000, 200000, True),        # cfg 3 size
    ],
)
@pytest.mark.parametrize("wide", [False, True])
def test_synthetic_trees_exact(gen, n_leaves, n_pairs, literal, wide):
    """fp32 edges in [0.5,1): fp64 path sums are exact, so the kernel must agree with
    the oracle bit for bit, not merely to 1e-12 -- in the compact 16-byte layout the
    library picks for such trees and in the wide (double-double) layout alike."""
    ft = gen(n_leaves, seed=1)
    T = SuchTree.from_flat(ft, _wide=wide)
    assert T.index_info["layout"] == (0 if wide else 1)
    ot = O.OracleTree(ft.parent, ft.distance)
    assert T.depth == ot.depth
    rng = np.random.default_rng(11)
    half = n_pairs // 2
    pairs = np.concatenate([
        rng.integers(0, ft.size, size=(half, 2)),             # any nodes
        2 * rng.integers(0, n_leaves, size=(n_pairs - half, 2)),  # leaves
    ]).astype(np.int64)
    # sprinkle near pairs (same block / same micro block paths) and a == b
    k = min(2000, n_pairs // 4)
    k -= k % 2
    pairs[:k, 1] = np.clip(pairs[:k, 0] + rng.integers(-40, 41, size=k), 0, ft.size - 1)
    pairs[k:k + 50, 1] = pairs[k:k + 50, 0]
    if literal:
        want_m = ot.mrca_bulk(pairs)
        want_d = ot.distances_f64(pairs)
    else:
        want_d, want_m = ot.distances_f64_climb(pairs, with_mrca=True)
    assert np.array_equal(T.common_ancestors_bulk(pairs), want_m)
    assert np.array_equal(T.distances_bulk(pairs), want_d)


def test_device_path_matches_host_path_and_generator():
    import torch

    ft = synth.yule_tree(50000, seed=2)
    T = SuchTree.from_flat(ft)
    n = 1000003  # odd: exercises the tail
    dev = torch.device("cuda", T.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for bits, dt in ((32, torch.int32), (64, torch.int64)):
        pairs = torch.empty((n + 1, 2), dtype=dt, device=dev)
        out = torch.empty(n + 2, dtype=torch.float64, device=dev)
        mr = torch.empty(n + 2, dtype=torch.int32, device=dev)
        T.random_leaf_pairs_device(1234, 0, n, pairs.data_ptr(), idx_bits=bits, stream=stream)
        T.distances_device(pairs.data_ptr(), n, out.data_ptr(), idx_bits=bits, d_mrca_ptr=mr.data_ptr(), stream=stream)
        T.check_range(stream)
        host_pairs = pairs[:n].cpu().numpy().astype(np.int64)
        assert np.array_equal(host_pairs, philox_ref.random_leaf_pairs(50000, 1234, 0, n))
        want = T.distances_bulk(host_pairs)
        assert np.array_equal(out[:n].cpu().numpy(), want)
        assert np.array_equal(mr[:n].cpu().numpy(), T.common_ancestors_bulk(host_pairs))
        # unaligned views take the scalar kernel: same bits
        p2 = pairs[1:]  # 8 or 16 bytes off
        o2 = out[1:]
        T.distances_device(p2.data_ptr(), n - 1, o2.data_ptr(), idx_bits=bits, stream=stream)
        T.check_range(stream)
        assert np.array_equal(o2[: n - 1].cpu().numpy(), want[1:n])
    # sharded generation reproduces the single stream
    a = torch.empty((1000, 2), dtype=torch.int32, device=dev)
    T.random_leaf_pairs_device(7, 0, 400, a.data_ptr(), stream=stream)
    T.random_leaf_pairs_device(7, 400, 600, a[400:].data_ptr(), stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(a.cpu().numpy(), philox_ref.random_leaf_pairs(50000, 7, 0, 1000))


def test_full_size_properties_cfg2():
    """BASELINE cfg 2 at full size (100k-leaf random tree, 1e8 pairs): properties that
    need no oracle -- symmetry d(a,b)=d(b,a), identity d(a,a)=0, agreement of the
    vector int32 kernel with the scalar int64 kernel, and the closed form
    rd[a]+rd[b]-2rd[mrca] evaluated in torch on exported root distances."""
    import torch

    ft = synth.yule_tree(100000, seed=1)
    T = SuchTree.from_flat(ft)
    n = 100_000_000
    dev = torch.device("cuda", T.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p32 = torch.empty((n, 2), dtype=torch.int32, device=dev)
    T.random_leaf_pairs_device(2, 0, n, p32.data_ptr(), stream=stream)
    d = torch.empty(n, dtype=torch.float64, device=dev)
    m = torch.empty(n, dtype=torch.int32, device=dev)
    T.distances_device(p32.data_ptr(), n, d.data_ptr(), d_mrca_ptr=m.data_ptr(), stream=stream)
    T.check_range(stream)
    # swapped pairs, int64, scalar path (offset view)
    p64 = torch.empty((n + 1, 2), dtype=torch.int64, device=dev)
    p64[1:, 0] = p32[:, 1]
    p64[1:, 1] = p32[:, 0]
    d2 = torch.empty(n + 1, dtype=torch.float64, device=dev)
    T.distances_device(p64[1:].data_ptr(), n, d2[1:].data_ptr(), idx_bits=64, stream=stream)
    T.check_range(stream)
    assert torch.equal(d, d2[1:])
    del p64, d2
    same = p32[:, 0] == p32[:, 1]
    assert bool((d[same] == 0).all()) and bool((d[~same] > 0).all())
    assert bool((m[same] == p32[:, 0][same]).all())
    depth, hi, lo = T.export_index()
    rd = torch.from_numpy(hi).to(dev)  # synthetic edges: lo == 0, sums exact
    assert not lo.any()
    a, b = p32[:, 0].long(), p32[:, 1].long()
    assert torch.equal(d, rd[a] + rd[b] - 2 * rd[m.long()])
    dep = torch.from_numpy(depth).to(dev)
    assert bool((dep[m.long()] <= torch.minimum(dep[a], dep[b])).all())
    assert bool(((m.long() >= torch.minimum(a, b)) & (m.long() <= torch.maximum(a, b))).all())
    assert bool((m[~same] % 2 == 1).all())  # MRCA of two distinct leaves is internal (odd id)


def test_error_contract():
    """test_new_api.py:841-866 plus the InvalidNodeError id rule of MuchTree.pyx:897-903."""
    from suchtree_b200 import InvalidNodeError, NodeNotFoundError

    T = SuchTree(os.path.join(GOLDEN, "data", "test.tree"))
    with pytest.raises(ValueError):
        T.distances_bulk(np.array([[1, 2, 3]]))
    with pytest.raises(ValueError):
        T.distances_bulk(np.array([[1, 2]], dtype=np.int32))
    with pytest.raises(ValueError):
        T.distances_bulk(np.zeros((0, 2), dtype=np.int64))
    with pytest.raises(InvalidNodeError) as e:
        T.distances_bulk(np.array([[0, 2], [0, 29], [4, 1000]], dtype=np.int64))
    assert e.value.node_id == 1000 and e.value.tree_size == 29
    with pytest.raises(InvalidNodeError) as e:
        T.distances_bulk(np.array([[0, 2], [-7, 3], [-2, 1]], dtype=np.int64))
    assert e.value.node_id == -7
    with pytest.raises(InvalidNodeError) as e:  # both: the max wins
        T.distances_bulk(np.array([[-5, 2], [0, 31]], dtype=np.int64))
    assert e.value.node_id == 31
    # the tree still answers afterwards
    assert T.distances_bulk(np.array([[0, 2]], dtype=np.int64))[0] > 0
    with pytest.raises(InvalidNodeError):
        T.distance(0, 29)
    with pytest.raises(NodeNotFoundError):
        T.distance("nope", "Ttal")
    with pytest.raises(TypeError):
        T.distance(1.5, 2)
    with pytest.raises(TypeError):
        T.distances_by_name(("Ttal", "Tbot"))
    with pytest.raises(TypeError):
        T.distances_by_name([("Ttal", 3)])
    with pytest.raises(NodeNotFoundError):
        T.distances_by_name([("Ttal", "nope")])
    # list input and strided views are accepted like the reference's memoryview
    assert T.distances_bulk([(0, 2), (4, 6)]).shape == (2,)
    big = np.arange(40, dtype=np.int64).reshape(10, 4) % 29
    view = big[::2, 1:4:2]
    assert not view.flags.c_contiguous
    assert np.array_equal(T.distances_bulk(view), T.distances_bulk(np.ascontiguousarray(view)))
    rev = np.ascontiguousarray(view)[::-1, ::-1]
    assert np.array_equal(T.distances_bulk(rev), T.distances_bulk(np.ascontiguousarray(rev)))


def test_mrca_properties():
    """SuchTree/tests/test_SuchTree.py:163-183, test_new_api.py:454-468."""
    from itertools import combinations

    T = SuchTree(os.path.join(GOLDEN, "data", "test.tree"))
    for a, b in combinations(T.leaves.values(), 2):
        m = T.common_ancestor(a, b)
        assert m == T.common_ancestor(b, a)
        with pytest.warns(DeprecationWarning):
            assert T.mrca(a, b) == m
        assert T.is_ancestor(m, a) == 1 and T.is_ancestor(m, b) == 1
        assert a in T.get_descendants(m) and b in T.get_descendants(m)
    for a, b in combinations(T.leaves.keys(), 2):
        assert T.common_ancestor(a, b) == T.common_ancestor(T.leaves[a], T.leaves[b])


def test_more_than_2_pow_32_pairs_in_one_launch():
    """64-bit indexing: the reference's `unsigned int` loop counters
    (MuchTree.pyx:923-930) wrap above 2^32 pairs per call.  One launch over 2^32 + 2^20
    device-resident pairs; the results beyond the 32-bit boundary must equal a separate
    small launch over the same pairs."""
    import torch

    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 32) + (1 << 20)
    if free < 17 * n + (4 << 30):
        pytest.skip("needs ~75 GB of free device memory")
    ft = synth.yule_tree(100000, seed=1)
    T = SuchTree.from_flat(ft)
    dev = torch.device("cuda", T.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    pairs = torch.empty((n, 2), dtype=torch.int32, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    T.random_leaf_pairs_device(5, 0, n, pairs.data_ptr(), stream=stream)
    T.distances_device(pairs.data_ptr(), n, out.data_ptr(), stream=stream)
    T.check_range(stream)
    for lo in (0, (1 << 31) - 512, (1 << 32) - 512, n - 4096):
        m = min(4096, n - lo)
        ref = torch.empty(m, dtype=torch.float64, device=dev)
        T.distances_device(pairs[lo:lo + m].data_ptr(), m, ref.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(out[lo:lo + m], ref), lo
        # and the generator itself did not wrap: same pairs as a shard started at `lo`
        chk = torch.empty((m, 2), dtype=torch.int32, device=dev)
        T.random_leaf_pairs_device(5, lo, m, chk.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(pairs[lo:lo + m], chk), lo
    assert bool(torch.isfinite(out[-(1 << 20):]).all())
    del pairs, out
    torch.cuda.empty_cache()
