"""Single-call latency of the drop-in entry points (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth
T = SuchTree.from_flat(synth.yule_tree(100000, seed=1, names=True))
G = SuchTree("tests/golden/data/test.tree")
for name, tree in (("100k-leaf", T), ("gopher", G)):
    ids = list(tree.leaves.values())
    a, b = ids[1], ids[-2]
    tree.distance(a, b)
    for n in (1, 256, 4096, 16384, 65536, 262144, 300000):
        pairs = np.array([[a, b]] * n, dtype=np.int64)
        tree.distances_bulk(pairs)
        t0 = time.perf_counter()
        reps = 100
        for _ in range(reps):
            tree.distances_bulk(pairs)
        dt = (time.perf_counter() - t0) / reps
        print("%-10s distances_bulk n=%-6d %.1f us/call" % (name, n, dt * 1e6), flush=True)
    t0 = time.perf_counter()
    for _ in range(300):
        tree.distance(a, b)
    print("%-10s distance(a,b)            %.1f us/call" % (name, (time.perf_counter() - t0) / 300 * 1e6))
    t0 = time.perf_counter()
    for _ in range(300):
        tree.common_ancestor(a, b)
    print("%-10s common_ancestor(a,b)     %.1f us/call" % (name, (time.perf_counter() - t0) / 300 * 1e6))
