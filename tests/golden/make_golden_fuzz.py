"""Golden vectors on seeded RANDOM trees (odd NEWICK syntax: polytomies, missing / zero /
negative / scientific lengths, quoting, comments, support labels), produced by the UNMODIFIED
reference in the authoring container:

    python tests/golden/make_golden_fuzz.py

Output: tests/golden/fuzz_trees.npz with, per tree k = 0..39,
    t<k>__newick    the text given to SuchTree()
    t<k>__parent    reference structure (get_parent of every node)
    t<k>__depth     SuchTree.depth
    t<k>__pairs     (200,2) random node ids (leaves and internal nodes)
    t<k>__mrca      common_ancestor of the first 50 pairs
    t<k>__distance  distances_bulk(pairs)            (the reference's fp32 accumulation)
    t<k>__q, t<k>__t  (100,4) random quartets and quartet_topologies_bulk of them (size >= 4)
The generator is tests/test_host_logic.py::_random_newick; every number comes from reference code.
"""
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(REPO, "oracle"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import ref_loader  # noqa: E402
from test_host_logic import _random_newick  # noqa: E402

warnings.simplefilter("ignore", DeprecationWarning)


def main():
    M = ref_loader.load_reference()
    assert M is not None, "reference not built"
    rng = random.Random(4242)
    nrng = np.random.default_rng(4243)
    out = {}
    for k in range(40):
        text = _random_newick(rng, rng.randint(2, 60))
        T = M.SuchTree(text)
        pairs = nrng.integers(0, T.size, size=(200, 2)).astype(np.int64)
        out["t%d__newick" % k] = np.array(text)
        out["t%d__parent" % k] = np.array([T.get_parent(i) for i in range(T.size)], dtype=np.int32)
        out["t%d__depth" % k] = np.int64(T.depth)
        out["t%d__pairs" % k] = pairs
        out["t%d__mrca" % k] = np.array([T.common_ancestor(int(a), int(b)) for a, b in pairs[:50]], dtype=np.int32)
        out["t%d__distance" % k] = T.distances_bulk(pairs)
        if T.size >= 4:
            q = nrng.integers(0, T.size, size=(100, 4)).astype(np.int64)
            out["t%d__q" % k] = q
            out["t%d__t" % k] = T.quartet_topologies_bulk(q)
    np.savez_compressed(os.path.join(HERE, "fuzz_trees.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
