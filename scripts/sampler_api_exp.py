"""Per-cycle cost of SuchLinkedTrees.sample_linked_distances (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, pandas as pd
from suchtree_b200 import SuchLinkedTrees, SuchTree, synth
A2 = SuchTree.from_flat(synth.yule_tree(14, seed=8, names=True))
B2 = SuchTree.from_flat(synth.yule_tree(103_446, seed=9, names=True))
rng2 = np.random.default_rng(10)
mat = np.zeros((14, 103_446), dtype=np.int8)
cols = rng2.choice(103_446, size=44_904, replace=False)
mat[rng2.integers(0, 14, size=44_904), cols] = 1
SLT = SuchLinkedTrees(A2, B2, pd.DataFrame(mat, index=list(A2.leaves.keys()), columns=list(B2.leaves.keys())))
SLT.sample_linked_distances(sigma=0.0, buckets=64, n=4096, maxcycles=1)
for cycles in (1, 10, 10, 40, 40):
    t0 = time.perf_counter()
    SLT.sample_linked_distances(sigma=0.0, buckets=64, n=4096, maxcycles=cycles)
    dt = time.perf_counter() - t0
    print(cycles, "cycles: %.2f ms per cycle, %.3e samples/s" % (1e3 * dt / cycles, cycles * 64 * 4096 / dt), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
SLT.sample_linked_distances(sigma=0.0, buckets=64, n=4096, maxcycles=20)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
