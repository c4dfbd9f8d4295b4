"""Golden vectors for the quartet-topology path, produced by the
UNMODIFIED reference in the authoring container:

    python tests/golden/make_golden_quartets.py

Output: tests/golden/quartets.npz with, per tree <name>,
    <name>__q   (n,4) int64 quartets: leaves in arbitrary order, plus rows with
                internal nodes and repeated ids (the reference's fall-through cases)
    <name>__t   (n,4) int64 = reference.quartet_topologies_bulk(<name>__q)
Every number comes from reference code (oracle/_ref, MuchTree.pyx:1271-1376).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)
import ref_loader  # noqa: E402

warnings.simplefilter("ignore", DeprecationWarning)
DATA = os.path.join(HERE, "data")


def main():
    M = ref_loader.load_reference()
    assert M is not None, "reference not built"
    with open(os.path.join(HERE, "trees.json")) as f:
        trees = json.load(f)
    rng = np.random.default_rng(20261018)
    out = {}
    sources = {name: rec.get("newick") or os.path.join(DATA, name) for name, rec in trees.items()}
    sources["big_ml"] = "/root/reference/data/bigtrees/ml.tree"
    sources["big_nj"] = "/root/reference/data/bigtrees/nj.tree"
    for name, src in sources.items():
        T = M.SuchTree(src)
        leaves = np.array(list(T.leaves.values()), dtype=np.int64)
        n = 3000 if name.startswith("big_") else 600
        parts = []
        if len(leaves) >= 4:
            # distinct leaves, arbitrary order
            q = np.stack([rng.permutation(leaves)[:4] for _ in range(n)]) if len(leaves) < 64 else \
                leaves[rng.integers(0, len(leaves), size=(n, 4))]
            parts.append(q)
        # any nodes (internal too), repeats allowed
        parts.append(rng.integers(0, T.size, size=(n // 2, 4)).astype(np.int64))
        # forced repeats: (a,a,b,c), (a,b,a,b), (a,a,a,a)
        r = rng.integers(0, T.size, size=(60, 4)).astype(np.int64)
        r[:20, 1] = r[:20, 0]
        r[20:40, 2] = r[20:40, 0]
        r[20:40, 3] = r[20:40, 1]
        r[40:, 1:] = r[40:, :1]
        parts.append(r)
        q = np.ascontiguousarray(np.concatenate(parts), dtype=np.int64)
        t = T.quartet_topologies_bulk(q)
        out[name + "__q"] = q
        out[name + "__t"] = np.asarray(t, dtype=np.int64)
        print(name, q.shape)
    np.savez_compressed(os.path.join(HERE, "quartets.npz"), **out)


if __name__ == "__main__":
    main()
