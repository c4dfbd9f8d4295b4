"""Per-clade scan timing (run under gpurun): the bench's clade workload -- two 100k-leaf Yule
trees, 100k random links, clade_pearson over every internal node of TreeB with 10..2500 links
(the filter of docs/examples/SuchLinkedTree_examples.md:289-296) -- and the same scan done the
way the reference does it, one subset_b + linked_pearson round trip per clade, on a sample of
the clades.
   python scripts/clade_exp.py [n_leaves]  -> gpurun_out/clade_exp.json"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from suchtree_b200 import SuchLinkedTrees, SuchTree, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    fa, fb = synth.yule_tree(n, seed=4, names=True), synth.yule_tree(n, seed=5, names=True)
    A, B = SuchTree.from_flat(fa), SuchTree.from_flat(fb)
    rng = np.random.default_rng(6)
    la, lb = 2 * rng.integers(0, n, n), 2 * rng.integers(0, n, n)
    t0 = time.perf_counter()
    SLT = SuchLinkedTrees.from_linklist(A, B, np.stack([lb, la], axis=1).astype(np.int64))
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = SLT.clade_pearson(min_links=10, max_links=2500)
    first_s = time.perf_counter() - t0
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        res = SLT.clade_pearson(min_links=10, max_links=2500)
        times.append(time.perf_counter() - t0)
    done = np.isfinite(res["r"])
    pairs = int(res["n_pairs"].sum())
    out = {
        "n_leaves": n, "n_links": int(SLT.n_links), "clades": int(res["node_ids"].shape[0]),
        "clades_computed": int(done.sum()), "link_pairs": pairs, "construction_s": build_s,
        "first_call_s": first_s, "s_per_scan": min(times), "link_pairs_per_s": pairs / min(times),
    }
    # the reference's way: one host round trip per clade (our own fused linked_pearson per clade)
    pick = np.nonzero(done)[0]
    pick = pick[:: max(1, len(pick) // 200)]
    t0 = time.perf_counter()
    worst = 0.0
    for k in pick:
        SLT.subset_b(int(res["node_ids"][k]))
        worst = max(worst, abs(SLT.linked_pearson() - res["r"][k]))
    loop_s = (time.perf_counter() - t0) / len(pick)
    out.update({"per_clade_loop_s_per_clade": loop_s, "per_clade_loop_s_extrapolated": loop_s * int(done.sum()),
                "per_clade_loop_sample": int(len(pick)), "max_abs_r_difference_on_sample": worst})
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "clade_exp.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
