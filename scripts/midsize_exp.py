"""Per-call time of distances_bulk for mid-size calls against the pipeline's chunk size
(run under gpurun)."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth
T = SuchTree.from_flat(synth.yule_tree(100000, seed=1))
rng = np.random.default_rng(0)
out = []
for n in (300_000, 1_000_000, 2_000_000, 4_000_000, 10_000_000, 30_000_000):
    P = 2 * rng.integers(0, 100000, size=(n, 2))
    row = {"n": n}
    for chunk in (0, 65536, 131072, 262144, 524288, 1048576, 2097152):
        if chunk:
            os.environ["SUCHTREE_B200_CHUNK_PAIRS"] = str(chunk)
        else:
            os.environ.pop("SUCHTREE_B200_CHUNK_PAIRS", None)
        for _ in range(3):
            T.distances_bulk(P)
        reps = max(5, min(50, 30_000_000 // n))
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); T.distances_bulk(P); ts.append(time.perf_counter() - t0)
        row[str(chunk)] = sorted(ts)[len(ts) // 2]
    out.append(row)
    print(n, {k: ("%.3f ms" % (v * 1e3)) for k, v in row.items() if k != "n"}, flush=True)
json.dump(out, open("gpurun_out/midsize.json", "w"), indent=1)
