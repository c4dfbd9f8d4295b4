"""Generates tests/golden/* by running the UNMODIFIED reference in the authoring
container (it cannot travel to the GPU box, the vectors can).

    python tests/golden/make_golden.py

Inputs : /root/reference (read-only) + oracle/_ref (reference extension compiled
         by oracle/build_ref.py, imported through oracle/ref_loader.py with the
         dendropy stand-in of oracle/newick_ref.py).
Outputs: tests/golden/data/   small input files of the reference's test-suite and
                              data/ directory (tree/links/matrix fixtures, the
                              two bigtrees gzip'ed), copied verbatim
         tests/golden/trees.json      per-tree structure as the reference built it
         tests/golden/pairs_*.npz     (pairs, mrca, distance) straight from the
                                      reference's distances_bulk / common_ancestor
         tests/golden/linked_*.npz    linklist, linked_distances(), pearson()
         tests/golden/sampled_*.npz   sample_linked_distances() under a fixed
                                      numpy seed (pins the xorshift64* stream)
         tests/golden/published.json  vectors published in the reference's docs
Every number here comes from reference code; nothing is computed by this repo's
product or oracle.
"""
import gzip
import json
import os
import shutil
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)
import ref_loader  # noqa: E402

warnings.simplefilter("ignore", DeprecationWarning)
REF = "/root/reference"
DATA = os.path.join(HERE, "data")

COPY = {
    "test.tree": "SuchTree/tests/test.tree",
    "lice.tree": "SuchTree/tests/lice.tree",
    "links.csv": "SuchTree/tests/links.csv",
    "test.matrix": "SuchTree/tests/test.matrix",
    "support_int.tree": "SuchTree/tests/support_int.tree",
    "support_float.tree": "SuchTree/tests/support_float.tree",
    "support_comment.tree": "SuchTree/tests/support_comment.tree",
    "fishworm_host.tree": "data/fish-worm/host.tree",
    "fishworm_guest.tree": "data/fish-worm/guest.tree",
    "fishworm_links.csv": "data/fish-worm/links.csv",
    "perfect0_host.tree": "data/simulated/perfect/perfect0/host.tree",
    "perfect0_guest.tree": "data/simulated/perfect/perfect0/guest.tree",
    "perfect0_links.csv": "data/simulated/perfect/perfect0/links.csv",
    "arr1_plant.tree": "data/plant-pollinators/arr1/plant.tree",
    "arr1_animal.tree": "data/plant-pollinators/arr1/animal.tree",
    "arr1_links.csv": "data/plant-pollinators/arr1/arr1_links.csv",
    "bigtrees_host.tree": "data/bigtrees/host.tree",
}
COPY_GZ = {"ml.tree.gz": "data/bigtrees/ml.tree", "nj.tree.gz": "data/bigtrees/nj.tree"}

# hand-written NEWICK strings exercising the id-defining rules
INLINE = {
    "abcd": "(A,B,(C,D));",
    "poly5": "(A:1,B:2,C:3,D:4,E:5);",
    "nested_poly": "((A:1,B:1,C:1,D:1):0.5,(E:2,F:2,G:2):0.25,H:3,I:1);",
    "zero_and_missing": "((A:0,B):1.5,(C:0.0,D:1e-3):0,E:2.5);",
    "quoted": "(('a b':1,'it''s':2):1,(c_d:1,'e(f)':1.0):2);",
    "comments": "((A[x]:1[c],B:2)[&&NHX:S=1]:1,(C:1,D:1)0.9[y]:2)root;",
    "negative": "((A:-0.5,B:1):0.25,(C:2,D:-0.00001):3);",
    "ladder": "((((((A:1,B:2):3,C:4):5,D:6):7,E:8):9,F:10):11,G:12);",
    "sci": "((A:1e-10,B:2.5E+3):1e5,(C:3.25e-7,D:4):5e-1);",
}


def tree_record(M, src):
    T = M.SuchTree(src)
    n = T.size
    par = [int(T.get_parent(i)) for i in range(n)]
    ch = [T.get_children(i) for i in range(n)]
    rec = dict(
        size=int(n),
        depth=int(T.depth),
        num_leaves=int(T.num_leaves),
        root=int(T.root_node),
        leaves={k: int(v) for k, v in T.leaves.items()},
        parent=par,
        left=[int(c[0]) for c in ch],
        right=[int(c[1]) for c in ch],
        # fp32 edge lengths as the reference stored them (exact via float.hex)
        distance_hex=[float(np.float32(d)).hex() for (_, _, d) in _edges(T, n)],
        support=[float(T.get_support(i)) for i in range(n)],
    )
    return T, rec


def _edges(T, n):
    # Node.distance is not exported directly; distance(node, parent) is the fp32 edge
    # (a single-term fp32 sum), root has the -1 sentinel.
    out = []
    for i in range(n):
        p = T.get_parent(i)
        out.append((i, p, -1.0 if p == -1 else T.distance(i, p)))
    return out


def pair_record(T, rng, n_pairs, leaves_only=False):
    n = T.size
    if leaves_only:
        ids = np.array(list(T.leaves.values()), dtype=np.int64)
        pairs = ids[rng.integers(0, len(ids), size=(n_pairs, 2))]
    else:
        pairs = rng.integers(0, n, size=(n_pairs, 2)).astype(np.int64)
    d = T.distances_bulk(pairs)
    m = np.array([T.common_ancestor(int(a), int(b)) for a, b in pairs], dtype=np.int32)
    return pairs, m, d


def main():
    M = ref_loader.load_reference()
    assert M is not None, "reference not built"
    os.makedirs(DATA, exist_ok=True)
    for dst, src in COPY.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(DATA, dst))
    for dst, src in COPY_GZ.items():
        with open(os.path.join(REF, src), "rb") as f, gzip.GzipFile(
            os.path.join(DATA, dst), "wb", mtime=0
        ) as g:
            g.write(f.read())

    rng = np.random.default_rng(20261017)
    trees = {}
    small_sources = {k: os.path.join(DATA, k) for k in COPY if k.endswith(".tree")}
    small_sources.update(INLINE)
    for name, src in small_sources.items():
        T, rec = tree_record(M, src)
        if name in INLINE:
            rec["newick"] = src
        trees[name] = rec
        n = T.size
        # all ordered node pairs for small trees (leaves AND internal nodes)
        if n <= 64:
            pairs = np.array([(a, b) for a in range(n) for b in range(n)], dtype=np.int64)
            d = T.distances_bulk(pairs)
            m = np.array([T.common_ancestor(int(a), int(b)) for a, b in pairs], dtype=np.int32)
        else:
            pairs, m, d = pair_record(T, rng, 4000)
        np.savez_compressed(os.path.join(HERE, "pairs_%s.npz" % name), pairs=pairs, mrca=m, distance=d)
    with open(os.path.join(HERE, "trees.json"), "w") as f:
        json.dump(trees, f, indent=0, sort_keys=True)

    # big trees: structure summary + sampled pairs (any node) + leaf pairs
    big = {}
    for name in ("ml", "nj"):
        T = M.SuchTree(os.path.join(REF, "data/bigtrees/%s.tree" % name))
        leaves = T.leaves
        names = list(leaves.keys())
        big[name] = dict(
            size=int(T.size), depth=int(T.depth), num_leaves=int(T.num_leaves), root=int(T.root_node),
            first_leaves={k: int(leaves[k]) for k in names[:50]},
            last_leaves={k: int(leaves[k]) for k in names[-50:]},
        )
        par = np.array([T.get_parent(i) for i in range(T.size)], dtype=np.int32)
        p1, m1, d1 = pair_record(T, rng, 20000)
        p2, m2, d2 = pair_record(T, rng, 20000, leaves_only=True)
        np.savez_compressed(
            os.path.join(HERE, "pairs_big_%s.npz" % name),
            parent=par, pairs=np.concatenate([p1, p2]), mrca=np.concatenate([m1, m2]),
            distance=np.concatenate([d1, d2]),
        )
    with open(os.path.join(HERE, "bigtrees.json"), "w") as f:
        json.dump(big, f, indent=0, sort_keys=True)

    # linked trees
    linked_sets = {
        "gopher_louse": ("test.tree", "lice.tree", "links.csv"),
        "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"),
        "perfect0": ("perfect0_host.tree", "perfect0_guest.tree", "perfect0_links.csv"),
        "arr1": ("arr1_plant.tree", "arr1_animal.tree", "arr1_links.csv"),
    }
    for name, (ta, tb, lk) in linked_sets.items():
        T1 = M.SuchTree(os.path.join(DATA, ta))
        T2 = M.SuchTree(os.path.join(DATA, tb))
        links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
        if set(links.index) != set(T1.leaves.keys()):
            links = links.T
        np.random.seed(12345)
        expected_seed = np.random.randint(np.iinfo(np.uint64).max >> 1)
        np.random.seed(12345)
        SLT = M.SuchLinkedTrees(T1, T2, links)
        res = SLT.linked_distances()
        r32 = M.pearson(res["TreeA"], res["TreeB"])
        out = dict(
            linklist=np.array(SLT.linklist), n_links=SLT.n_links,
            TreeA=res["TreeA"], TreeB=res["TreeB"], ids_A=res["ids_A"], ids_B=res["ids_B"],
            n_pairs=res["n_pairs"], pearson=r32, rng_seed=np.uint64(expected_seed),
        )
        np.savez_compressed(os.path.join(HERE, "linked_%s.npz" % name), **out)
        # sampled distances: small buckets so that the vectors stay small
        s = SLT.sample_linked_distances(sigma=0.05, buckets=8, n=64, maxcycles=50)
        if s is not None:
            np.savez_compressed(
                os.path.join(HERE, "sampled_%s.npz" % name),
                TreeA=s["TreeA"], TreeB=s["TreeB"], n_pairs=s["n_pairs"], n_samples=s["n_samples"],
                deviation_a=s["deviation_a"], deviation_b=s["deviation_b"],
                sigma=0.05, buckets=8, n=64, maxcycles=50, rng_seed=np.uint64(expected_seed),
            )
        else:
            print("sampled", name, "did not converge; no fixture")

    # pearson() on assorted vectors, straight from the reference.  Only n and r are
    # stored; tests regenerate x, y from the same numpy Generator stream (seed 7).
    pv = {"n": [], "r": []}
    r2 = np.random.default_rng(7)
    for n in [2, 3, 17, 1000, 100000]:
        x = r2.random(n)
        y = 0.3 * x + r2.random(n)
        pv["n"].append(n)
        pv["r"].append(float(M.pearson(x, y)).hex())
    with open(os.path.join(HERE, "pearson_vectors.json"), "w") as f:
        json.dump(pv, f)

    published = {
        "_source": "numbers printed in the reference's own docs/notebooks (no code of ours involved)",
        "gopher_louse_pearsonr": {"value": 0.49018498968585178, "src": "data/gopher-louse/Gopher-Louse.ipynb cell 9"},
        "gopher_louse_kendalltau": {"value": 0.20975684102929301, "src": "data/gopher-louse/Gopher-Louse.ipynb cell 9"},
        "gopher_ids": {"Ttal": 26, "Tbot": 28, "src": "data/gopher-louse/Gopher-Louse.ipynb:1238-1263"},
        "lice_ids": {"Tbar": 30, "Tmin": 32, "src": "data/gopher-louse/Gopher-Louse.ipynb:1238-1263"},
        "gopher_louse_linklist": {
            "value": [[10, 4], [16, 10], [32, 28], [22, 18], [26, 22], [0, 26], [30, 26], [20, 14], [2, 28],
                      [4, 12], [8, 2], [28, 24], [12, 6], [24, 20], [14, 8], [18, 16], [6, 0]],
            "src": "data/gopher-louse/Gopher-Louse.ipynb:1238-1263 (SLT.linklist, rows = [TreeB id, TreeA id])",
        },
        "bigtrees_host_leaves": {
            "value": {
                "Tropheus_moorii": 0, "Lobochilotes_labiatus": 2, "Tanganicodus_irsacae": 4,
                "Cyprichromis_coloratus": 6, "Haplotaxodon_microlepis": 8, "Perissodus_microlepis": 10,
                "Plecodus_straeleni": 12, "Xenotilapia_flavipinnis": 14, "Triglachromis_otostigma": 16,
                "Reganochromis_calliurus": 18, "Trematochromis_benthicola": 20,
                "Lepidiolamprologus_profundicola": 22, "Neolamprologus_buescheri": 24,
                "Chalinochromis_brichardi": 26,
            },
            "src": "docs/examples/SuchTree_examples.md:99-112 (T.leaves of data/bigtrees/host.tree)",
        },
    }
    with open(os.path.join(HERE, "published.json"), "w") as f:
        json.dump(published, f, indent=1, sort_keys=True)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
