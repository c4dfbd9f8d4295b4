"""Joined link records (one 32-byte record per link: both trees' root distances and keys) against
separate link rows + node records, for the Philox sampler, the exhaustive moments and the clade
scan; results compared bit for bit.  Run under gpurun -> gpurun_out/joined_exp.json"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import SuchLinkedTrees, SuchTree, synth

def moments_tuple(m):
    return (m.n, m.sx, m.sy, m.sxx, m.syy, m.sxy)

def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return r, min(ts)

out = {}
n = 100_000
A, B = SuchTree.from_flat(synth.yule_tree(n, seed=4)), SuchTree.from_flat(synth.yule_tree(n, seed=5))
rng = np.random.default_rng(6)
la, lb = 2 * rng.integers(0, n, n), 2 * rng.integers(0, n, n)
ll = np.stack([lb, la], axis=1).astype(np.int64)
res = {}
for mode in ("0", "1"):
    os.environ["SUCHTREE_B200_JOINED"] = mode
    SLT = SuchLinkedTrees.from_linklist(A, B, ll)
    ns = 125_000_000
    m, t = timed(lambda: SLT.sample_moments(ns, seed=7, x0=20.0, y0=20.0))
    res["sampler", mode] = (moments_tuple(m), ns / t)
    SLT2 = SuchLinkedTrees.from_linklist(A, B, ll[:44904])
    npairs = 44904 * 44903 // 2
    m, t = timed(lambda: SLT2.linked_moments(0, npairs, x0=20.0, y0=20.0))
    res["exhaustive", mode] = (moments_tuple(m), npairs / t)
    r, t = timed(lambda: SLT.clade_pearson(min_links=10, max_links=2500), reps=3)
    res["clade", mode] = ((int(r["n_pairs"].sum()), float(np.nansum(r["r"]))), int(r["n_pairs"].sum()) / t)
    if mode == "1":
        rr = r
    else:
        r0 = r
for what in ("sampler", "exhaustive", "clade"):
    a, b = res[what, "0"], res[what, "1"]
    out[what] = {"separate_per_s": a[1], "joined_per_s": b[1], "speedup": b[1] / a[1], "bit_identical": a[0] == b[0]}
    print(what, out[what], flush=True)
out["clade"]["r_identical"] = bool(np.array_equal(r0["r"], rr["r"], equal_nan=True))
print(out["clade"]["r_identical"])
json.dump(out, open("gpurun_out/joined_exp.json", "w"), indent=1)
