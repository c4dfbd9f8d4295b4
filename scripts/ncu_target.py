"""A short, single-kernel workload for `ncu --set full` captures (run under gpurun):
    python scripts/ncu_target.py pairs  <yule|balanced|caterpillar> <n_pairs> [paired]
    python scripts/ncu_target.py quartets <n_quartets> <qpt> <idx_bits>
    python scripts/ncu_target.py linked      (cfg 4: 3 sampler launches, then 3 exhaustive-moment launches)
3 warm-up launches + 2 launches of the kernel under study on device-resident input."""
import os, sys
sys.path.insert(0, '.')
import torch
from suchtree_b200 import SuchTree, synth

what = sys.argv[1]
dev = torch.device('cuda', 0)
s = torch.cuda.current_stream(dev).cuda_stream
if what == 'pairs':
    shape, n = sys.argv[2], int(sys.argv[3])
    if len(sys.argv) > 4:
        os.environ['SUCHTREE_B200_PAIRED'] = '1' if sys.argv[4] == 'paired' else '0'
    ft = {'yule': lambda: synth.yule_tree(100000, seed=1), 'balanced': lambda: synth.balanced_tree(1_000_000, seed=3),
          'caterpillar': lambda: synth.caterpillar_tree(1_000_000, seed=3)}[shape]()
    T = SuchTree.from_flat(ft, device=0)
    pairs = torch.empty((n, 2), dtype=torch.int32, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    T.random_leaf_pairs_device(3, 0, n, pairs.data_ptr(), idx_bits=32, stream=s)
    for _ in range(5):
        T.distances_device(pairs.data_ptr(), n, out.data_ptr(), idx_bits=32, stream=s)
elif what == 'linked':
    import numpy as np
    from suchtree_b200 import SuchLinkedTrees
    n = 100_000
    A, B = SuchTree.from_flat(synth.yule_tree(n, seed=4)), SuchTree.from_flat(synth.yule_tree(n, seed=5))
    rng = np.random.default_rng(6)
    ll = np.stack([2 * rng.integers(0, n, n), 2 * rng.integers(0, n, n)], axis=1).astype(np.int64)
    S = SuchLinkedTrees.from_linklist(A, B, ll)
    for _ in range(3):
        S.sample_moments(125_000_000, seed=7, x0=20.0, y0=20.0)
    S2 = SuchLinkedTrees.from_linklist(A, B, ll[:44904])
    for _ in range(3):
        S2.linked_moments(0, 126_020_270, x0=20.0, y0=20.0)
    T = A
else:
    n, qpt, bits = int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    os.environ['SUCHTREE_B200_QPT'] = qpt
    T = SuchTree.from_flat(synth.yule_tree(100000, seed=1), device=0)
    g = torch.Generator(device=dev).manual_seed(21)
    q = 2 * torch.randint(0, 100000, (n, 4), generator=g, device=dev, dtype=torch.int64)
    if bits == 32:
        q = q.to(torch.int32)
    o = torch.empty_like(q)
    for _ in range(5):
        T.quartet_topologies_device(q.data_ptr(), n, o.data_ptr(), stream=s, idx_bits=bits)
torch.cuda.synchronize()
T.check_range(s)
print('done')
