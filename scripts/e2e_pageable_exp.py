"""e2e throughput of distances_bulk with ordinary (pageable) numpy arrays (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth
T = SuchTree.from_flat(synth.yule_tree(100000, seed=1))
n = 100_000_000
P = 2 * np.random.default_rng(0).integers(0, 100000, size=(n, 2))
T.distances_bulk(P[:1000000])
for rep in range(3):
    t0 = time.perf_counter(); r = T.distances_bulk(P); dt = time.perf_counter() - t0
    print('pageable in, fresh pageable out: %.3e pairs/s' % (n / dt), flush=True)
out = np.empty(n)
for rep in range(3):
    t0 = time.perf_counter(); T.distances_bulk(P, out=out); dt = time.perf_counter() - t0
    print('pageable in, reused pageable out: %.3e pairs/s' % (n / dt), flush=True)
assert np.array_equal(out, r)
