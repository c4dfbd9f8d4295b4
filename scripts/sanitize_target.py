"""Every kernel of libsuchtree_b200.so once, on small inputs, for compute-sanitizer
(run under gpurun):   compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_target.py
Results are compared with the oracle so that a silent wrong answer also fails."""
import ctypes as C, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import oracle as O
from suchtree_b200 import SuchLinkedTrees, SuchTree, _lib, pearson, synth
from suchtree_b200.linked import moments_pearson

rng = np.random.default_rng(0)
for gen, kw in ((lambda: synth.yule_tree(3000, seed=3), {}), (lambda: synth.yule_tree(3000, seed=3), {'_block_shift': 3}),
                (lambda: synth.caterpillar_tree(2000, seed=3), {}), (lambda: synth.yule_tree(700, seed=4), {'_wide': True})):
    ft = gen()
    T = SuchTree.from_flat(ft, device=0, **kw)
    ot = O.OracleTree(ft.parent, ft.distance)
    for n in (1, 5, 4097, 300_001):  # tiny / small / medium / chunked host routes
        p = rng.integers(0, ft.size, size=(n, 2)).astype(np.int64)
        for paired in ('0', '1'):
            os.environ['SUCHTREE_B200_PAIRED'] = paired
            want, wm = ot.distances_f64_climb(p, with_mrca=True)
            assert np.array_equal(T.distances_bulk(p), want)
            assert np.array_equal(T.common_ancestors_bulk(p), wm)
        assert np.array_equal(T.distances_bulk(p[::-1][:, ::-1]), want[::-1])  # strided
    for qpt in ('1', '2'):
        os.environ['SUCHTREE_B200_QPT'] = qpt
        q = rng.integers(0, ft.size, size=(70_001, 4)).astype(np.int64)
        q[::5, 1] = q[::5, 0]
        assert np.array_equal(T.quartet_topologies_bulk(q), ot.quartet_topologies(q))
    leaves = list(range(0, 2 * min(ft.n_leaves, 700), 2))
    D = T.pairwise_distances(leaves)
    a, b = np.meshgrid(leaves, leaves, indexing='ij')
    assert np.array_equal(D, ot.distances_f64_climb(np.stack([a.ravel(), b.ravel()], 1).astype(np.int64)).reshape(D.shape))
    ids = rng.integers(0, ft.size, 300).tolist()
    D = T.pairwise_distances(ids)
    a, b = np.meshgrid(ids, ids, indexing='ij')
    assert np.array_equal(D, ot.distances_f64_climb(np.stack([a.ravel(), b.ravel()], 1).astype(np.int64)).reshape(D.shape))
    assert T.pairwise_distances().shape == (ft.n_leaves, ft.n_leaves)
    print('tree ok', ft.size, kw, flush=True)

fa, fb = synth.yule_tree(800, seed=31, names=True), synth.yule_tree(900, seed=32, names=True)
TA, TB = SuchTree.from_flat(fa, device=0), SuchTree.from_flat(fb, device=0)
ll = np.stack([2 * rng.integers(0, 900, 600), 2 * rng.integers(0, 800, 600)], axis=1).astype(np.int64)
S = SuchLinkedTrees.from_linklist(TA, TB, ll)
ld = S.linked_distances()
pa = O.linked_pairs(S.linklist)
assert np.array_equal(ld['ids_A'], pa[0]) and np.array_equal(ld['ids_B'], pa[1])
np.random.seed(1)
r = S.sample_linked_distances(sigma=0.05, buckets=8, n=256, maxcycles=20)
x = pearson(ld['TreeA'], ld['TreeB'])
assert abs(x - S.linked_pearson()) < 1e-9
assert abs(S.sample_pearson(100001, seed=3)) <= 1.0
scan = S.clade_pearson(min_links=3)
assert np.isfinite(scan['r']).any()
print('linked ok', flush=True)
