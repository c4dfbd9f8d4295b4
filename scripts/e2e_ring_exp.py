"""e2e of SuchTree.distances_bulk (pageable numpy in, fresh array out) against the staging
regime of the host pipeline: chunk size, ring depth, streaming vs ordinary stores, pack
fraction.  Run under gpurun; writes gpurun_out/e2e_ring.json."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchTree, synth

T = SuchTree.from_flat(synth.yule_tree(100000, seed=1))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
P = 2 * np.random.default_rng(0).integers(0, 100000, size=(n, 2))
KEYS = ("SUCHTREE_B200_CHUNK_PAIRS", "SUCHTREE_B200_RING", "SUCHTREE_B200_PACK_NT", "SUCHTREE_B200_PACK_FRACTION")

def run(cfg, reps=5):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ["SUCHTREE_B200_" + k] = str(v)
    r = T.distances_bulk(P)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = T.distances_bulk(P); ts.append(time.perf_counter() - t0)
    return r, n / min(ts), n / sorted(ts)[len(ts) // 2]

t0 = time.perf_counter(); ref = T.distances_bulk(P); first = time.perf_counter() - t0
print("first call %.3f s" % first, flush=True)
T.distances_bulk(P)  # second sighting: the input is registered here
out = {"first_call_s": first, "runs": []}
cfgs = [{"CHUNK_PAIRS": 0, "PACK_FRACTION": 0.45}, {"CHUNK_PAIRS": 0, "PACK_FRACTION": 1.0}]
for frac in (1.0, 0.8, 0.6):
    for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
        for ring in (3, 6, 12):
            if chunk * ((ring + 2) // 3) > (1 << 22):
                continue
            cfgs.append({"CHUNK_PAIRS": chunk, "RING": ring, "PACK_FRACTION": frac})
cfgs += [{"CHUNK_PAIRS": 1 << 20, "RING": 6, "PACK_FRACTION": 1.0, "PACK_NT": 1},
         {"CHUNK_PAIRS": 1 << 19, "RING": 6, "PACK_FRACTION": 1.0, "PACK_NT": 1},
         {"CHUNK_PAIRS": 0, "PACK_FRACTION": 1.0, "PACK_NT": 0}]
for cfg in cfgs:
    r, best, med = run(cfg)
    ok = bool(np.array_equal(r, ref))
    out["runs"].append({"cfg": cfg, "best": best, "median": med, "ok": ok})
    print(cfg, "best %.3e median %.3e" % (best, med), ok, flush=True)
json.dump(out, open("gpurun_out/e2e_ring.json", "w"), indent=1)
