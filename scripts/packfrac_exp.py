"""e2e of distances_bulk (1e8 pairs, numpy in, fresh array out) against the fraction of each chunk
whose ids are bit-packed on the host (the rest is DMA'd as int64 from the registered input).
Run under gpurun -> gpurun_out/packfrac_bits.json"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth
out = {}
for leaves, name in ((100_000, "yule100k"),):
    T = SuchTree.from_flat(synth.yule_tree(leaves, seed=1) if leaves == 100_000 else synth.balanced_tree(leaves, seed=3))
    n = 100_000_000
    P = 2 * np.random.default_rng(0).integers(0, leaves, size=(n, 2))
    ref = T.distances_bulk(P); T.distances_bulk(P)
    for frac in ("0.0", "0.3", "0.45", "0.6", "0.7", "0.8", "1.0"):
        os.environ["SUCHTREE_B200_PACK_FRACTION"] = frac
        r = T.distances_bulk(P)
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); r = T.distances_bulk(P); ts.append(time.perf_counter() - t0)
        ok = bool(np.array_equal(r, ref))
        out[name + "_" + frac] = {"best": n / min(ts), "median": n / sorted(ts)[3], "ok": ok}
        print(name, frac, "best %.3e median %.3e" % (n / min(ts), n / sorted(ts)[3]), ok, flush=True)
    del os.environ["SUCHTREE_B200_PACK_FRACTION"]
    del T, P, ref, r
json.dump(out, open("gpurun_out/packfrac_bits.json", "w"), indent=1)
