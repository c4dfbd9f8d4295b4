"""Golden vectors for the per-clade scan (SuchLinkedTrees.clade_pearson), produced by
running the reference's own scan loop -- the one of its co-phylogeny example,
docs/examples/SuchLinkedTree_examples.md:286-310 -- with the UNMODIFIED reference in the
authoring container:

    python tests/golden/make_golden_clades.py

Output: tests/golden/clades.npz with, per linked set <name> and scan <s> in
    b      for node in internal nodes of TreeB: SLT.subset_b(node)   (subset_a = all)
    a      for node in internal nodes of TreeA: SLT.subset_a(node)   (subset_b = all)
    ab     SLT.subset_a(<fixed internal node>) first, then the b scan
the arrays
    <name>__<s>__nodes     node ids scanned
    <name>__<s>__n_leafs   SLT.subset_b_size / subset_a_size after the call
    <name>__<s>__n_links   SLT.subset_n_links
    <name>__<s>__r32       reference pearson(ld['TreeA'], ld['TreeB'])   (fp32 accumulators, :62-79)
    <name>__<s>__r64       numpy fp64 correlation of the same two reference vectors (nan if undefined)
    <name>__ab__anode      the TreeA node of the `ab` scan
Every distance comes from reference code (oracle/_ref).
"""
import os
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
DATA = os.path.join(HERE, "data")

LINKED_SETS = {
    "gopher_louse": ("test.tree", "lice.tree", "links.csv"),
    "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv"),
    "perfect0": ("perfect0_host.tree", "perfect0_guest.tree", "perfect0_links.csv"),
    "arr1": ("arr1_plant.tree", "arr1_animal.tree", "arr1_links.csv"),
}


def _r64(x, y):
    if len(x) < 2:
        return np.nan
    x = np.asarray(x, np.float64) - np.mean(x)
    y = np.asarray(y, np.float64) - np.mean(y)
    den = np.sqrt(np.dot(x, x) * np.dot(y, y))
    return float(np.dot(x, y) / den) if den > 0 else np.nan


def scan(M, SLT, tree, subset, size_attr):
    nodes = [int(n) for n in tree.get_internal_nodes()]
    rec = {"nodes": [], "n_leafs": [], "n_links": [], "r32": [], "r64": []}
    for node in nodes:
        subset(node)
        if SLT.subset_n_links < 2:
            # no link pair: the reference's linked_distances() raises on the empty id
            # array (distances_bulk takes its max, :897); record the counts only
            ld = {"TreeA": np.empty(0), "TreeB": np.empty(0)}
        else:
            ld = SLT.linked_distances()
        rec["nodes"].append(node)
        rec["n_leafs"].append(int(getattr(SLT, size_attr)))
        rec["n_links"].append(int(SLT.subset_n_links))
        rec["r32"].append(float(M.pearson(ld["TreeA"], ld["TreeB"])) if len(ld["TreeA"]) else np.nan)
        rec["r64"].append(_r64(ld["TreeA"], ld["TreeB"]))
    return rec


def main():
    M = ref_loader.load_reference()
    assert M is not None, "reference not built"
    out = {}
    for name, (ta, tb, lk) in LINKED_SETS.items():
        T1 = M.SuchTree(os.path.join(DATA, ta))
        T2 = M.SuchTree(os.path.join(DATA, tb))
        links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
        if set(links.index) != set(T1.leaves.keys()):
            links = links.T
        SLT = M.SuchLinkedTrees(T1, T2, links)
        recs = {"b": scan(M, SLT, T2, SLT.subset_b, "subset_b_size")}
        SLT.subset_b(T2.root_node)
        recs["a"] = scan(M, SLT, T1, SLT.subset_a, "subset_a_size")
        SLT.subset_a(T1.root_node)
        # a fixed TreeA clade holding roughly half of the leaves, then the b scan under it
        inner = [int(n) for n in T1.get_internal_nodes()]
        anode = min(inner, key=lambda n: abs(len(T1.get_leaves(n)) - T1.num_leaves / 2.0))
        SLT.subset_a(anode)
        recs["ab"] = scan(M, SLT, T2, SLT.subset_b, "subset_b_size")
        out[name + "__ab__anode"] = np.int64(anode)
        for s, rec in recs.items():
            out["%s__%s__nodes" % (name, s)] = np.array(rec["nodes"], dtype=np.int64)
            out["%s__%s__n_leafs" % (name, s)] = np.array(rec["n_leafs"], dtype=np.int64)
            out["%s__%s__n_links" % (name, s)] = np.array(rec["n_links"], dtype=np.int64)
            out["%s__%s__r32" % (name, s)] = np.array(rec["r32"], dtype=np.float64)
            out["%s__%s__r64" % (name, s)] = np.array(rec["r64"], dtype=np.float64)
            print(name, s, len(rec["nodes"]), "clades, links", min(rec["n_links"]), "..", max(rec["n_links"]))
    np.savez_compressed(os.path.join(HERE, "clades.npz"), **out)


if __name__ == "__main__":
    main()
