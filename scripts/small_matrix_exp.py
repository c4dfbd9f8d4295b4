"""Latency of small host-facing calls (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth
G = SuchTree("tests/golden/data/test.tree")
T = SuchTree.from_flat(synth.yule_tree(100000, seed=1, names=True))
def t(fn, reps=200):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return 1e6 * (time.perf_counter() - t0) / reps
ids = list(range(0, 400, 2))
print("gopher pairwise_distances() all 15 leaves   %.0f us" % t(lambda: G.pairwise_distances()))
print("gopher nearest_neighbors(leaf, k=3)          %.0f us" % t(lambda: G.nearest_neighbors(0, k=3)))
print("100k tree pairwise_distances(200 ids)        %.0f us" % t(lambda: T.pairwise_distances(ids)))
print("100k tree pairwise_distances(2000 ids)       %.0f us" % t(lambda: T.pairwise_distances(list(range(0, 4000, 2))), 50))
print("100k tree distance(a, b)                     %.0f us" % t(lambda: T.distance(0, 19998)))
