"""Quartet topologies (MuchTree.pyx:1203-1421): the oracle against
vectors produced by the unmodified reference (CPU), and the CUDA path against both
(GPU, through the C ABI)."""
import os
from itertools import combinations

import numpy as np
import pytest
from conftest import DATA, GOLDEN, tree_source

import oracle as O
from suchtree_b200 import SuchTree, synth
from suchtree_b200.exceptions import InvalidNodeError, NodeNotFoundError


def _golden():
    return np.load(os.path.join(GOLDEN, "quartets.npz"))


def _sets(rows):
    return [frozenset((frozenset((int(a), int(b))), frozenset((int(c), int(d))))) for a, b, c, d in rows]


# ------------------------------------------------------------------- CPU -----
def test_oracle_quartets_match_reference(golden_trees):
    z = _golden()
    for name, rec in golden_trees.items():
        ot = O.OracleTree.from_newick(tree_source(name, rec))
        assert np.array_equal(ot.quartet_topologies(z[name + "__q"]), z[name + "__t"]), name


@pytest.mark.parametrize("name", ["ml", "nj"])
def test_oracle_quartets_big_trees(name):
    z = _golden()
    par = np.load(os.path.join(GOLDEN, "pairs_big_%s.npz" % name))["parent"]
    ot = O.OracleTree(par, np.zeros(par.shape[0], np.float32))
    assert np.array_equal(ot.quartet_topologies(z["big_%s__q" % name]), z["big_%s__t" % name])


# ------------------------------------------------------------------- GPU -----
@pytest.mark.gpu
@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("geom", [(0, 0), (2, 1), (4, 2)])
def test_gpu_quartets_small_trees(golden_trees, geom, wide):
    z = _golden()
    for name, rec in golden_trees.items():
        T = SuchTree(tree_source(name, rec), _block_shift=geom[0], _micro_shift=geom[1], _wide=wide)
        got = T.quartet_topologies_bulk(z[name + "__q"])
        assert got.dtype == np.int64 and np.array_equal(got, z[name + "__t"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ml", "nj"])
def test_gpu_quartets_big_trees(name):
    z = _golden()
    T = SuchTree(os.path.join(DATA, "%s.tree.gz" % name))
    assert np.array_equal(T.quartet_topologies_bulk(z["big_%s__q" % name]), z["big_%s__t" % name])


@pytest.mark.gpu
@pytest.mark.parametrize("gen,n_leaves,n_q", [(synth.yule_tree, 100000, 20000), (synth.caterpillar_tree, 3000, 3000),
                                              (synth.balanced_tree, 1 << 15, 20000)])
def test_gpu_quartets_synthetic_vs_oracle(gen, n_leaves, n_q):
    ft = gen(n_leaves, seed=9)
    T = SuchTree.from_flat(ft)
    ot = O.OracleTree(ft.parent, ft.distance)
    rng = np.random.default_rng(10)
    q = np.concatenate([2 * rng.integers(0, n_leaves, size=(n_q, 4)), rng.integers(0, ft.size, size=(n_q // 4, 4))])
    q = q.astype(np.int64)
    assert np.array_equal(T.quartet_topologies_bulk(q), ot.quartet_topologies(q))
    # strided input (the reference's memoryview takes any strides), odd length
    wide_arr = np.zeros((q.shape[0], 8), dtype=np.int64)
    wide_arr[:, ::2] = q
    view = wide_arr[:2001, ::2]
    assert np.array_equal(T.quartet_topologies_bulk(view), ot.quartet_topologies(np.ascontiguousarray(view)))


@pytest.mark.gpu
def test_gpu_quartets_device_path_and_full_size_properties():
    """1e7 quartets on the cfg-2 tree, device resident: every output row is a
    permutation of its input row, and the topology is invariant under permuting the
    input (a size-independent property; the oracle checks a sample)."""
    import torch

    ft = synth.yule_tree(100000, seed=1)
    T = SuchTree.from_flat(ft)
    dev = torch.device("cuda", T.device)
    n = 10_000_000
    g = torch.Generator(device=dev).manual_seed(5)
    # four distinct leaves per row: ranks a < b < c < d from sorted distinct draws
    r = torch.randint(0, 100000, (n, 4), generator=g, device=dev, dtype=torch.int64)
    r = torch.sort(r, dim=1).values
    keep = (r[:, 1:] != r[:, :-1]).all(dim=1)
    q = (2 * r[keep]).contiguous()
    n = q.shape[0]
    out = torch.empty_like(q)
    s = torch.cuda.current_stream(dev).cuda_stream
    T.quartet_topologies_device(q.data_ptr(), n, out.data_ptr(), stream=s)
    perm = torch.tensor([2, 0, 3, 1], device=dev)
    q2 = q[:, perm].contiguous()
    out2 = torch.empty_like(q2)
    T.quartet_topologies_device(q2.data_ptr(), n, out2.data_ptr(), stream=s)
    T.check_range(s)
    assert torch.equal(torch.sort(out, dim=1).values, q)

    def canon(t):  # unordered pair of unordered pairs -> canonical (n,4)
        p0 = torch.sort(t[:, :2], dim=1).values
        p1 = torch.sort(t[:, 2:], dim=1).values
        swap = (p0[:, 0] > p1[:, 0]).unsqueeze(1)
        return torch.where(swap, torch.cat([p1, p0], 1), torch.cat([p0, p1], 1))

    assert torch.equal(canon(out), canon(out2))
    ot = O.OracleTree(ft.parent, ft.distance)
    sample = q[:5000].cpu().numpy()
    assert np.array_equal(out[:5000].cpu().numpy(), ot.quartet_topologies(sample))


@pytest.mark.gpu
def test_gpu_quartet_api_mirrors_reference():
    """test_SuchTree.py:198-211, test_new_api.py:496-535 and the error contract."""
    T = SuchTree("(A,B,(C,D));")
    want = frozenset((frozenset(("A", "B")), frozenset(("C", "D"))))
    assert T.quartet_topology("A", "B", "C", "D") == want
    with pytest.warns(DeprecationWarning):
        assert T.get_quartet_topology("A", "B", "C", "D") == want
    T = SuchTree(os.path.join(DATA, "test.tree"))
    Q = np.array(list(combinations(T.leaves.values(), 4)))
    with pytest.warns(DeprecationWarning):
        q = T.quartet_topologies(Q)
    assert q.shape == Q.shape
    for a, b, c, d in q[::7]:
        assert T.quartet_topology(int(d), int(c), int(b), int(a)) == frozenset(
            (frozenset((int(a), int(b))), frozenset((int(c), int(d)))))
    names = list(T.leaves.keys())
    quartets = [tuple(c) for c in combinations(names, 4)][:3]
    tops = T.quartet_topologies_by_name(quartets)
    assert len(tops) == 3 and all(isinstance(t, frozenset) and len(t) == 2 for t in tops)
    ids = np.array([[T.leaves[x] for x in qd] for qd in quartets], dtype=np.int64)
    assert _sets(T.quartet_topologies_bulk(ids)) == [
        frozenset(frozenset(T.leaves[x] for x in s) for s in t) for t in tops]
    with pytest.raises(ValueError):
        T.quartet_topologies_bulk(np.zeros((3, 3), dtype=np.int64))
    with pytest.raises(InvalidNodeError) as e:
        T.quartet_topologies_bulk(np.array([[0, 2, 4, T.size]], dtype=np.int64))
    assert e.value.node_id == T.size
    with pytest.raises(InvalidNodeError) as e:
        T.quartet_topologies_bulk(np.array([[0, 2, -4, 6]], dtype=np.int64))
    assert e.value.node_id == -4
    with pytest.raises(NodeNotFoundError):
        T.quartet_topologies_by_name([("nope", names[0], names[1], names[2])])
    with pytest.raises(TypeError):
        T.quartet_topologies_by_name([(1, names[0], names[1], names[2])])
