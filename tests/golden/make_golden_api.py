"""Golden vectors for the non-accelerated part of the API (suchtree_b200/extras.py),
produced by the UNMODIFIED reference:

    python tests/golden/make_golden_api.py  ->  tests/golden/api.json

Per small tree: ancestors, support, get_nodes / get_internal_nodes, node tests,
bipartitions, paths, all traversals, adjacency / Laplacian / incidence / degree data,
networkx node and edge records, NEWICK text.  Per linked pair: the joint adjacency and
Laplacian matrices and the Laplacian spectrum.  Floats are stored as repr() strings of the
Python floats the reference returns, so that fp32-widened values compare exactly.
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


def fl(x):
    return repr(float(x))


def mat(a):
    a = np.asarray(a)
    return {"shape": list(a.shape), "values": [fl(v) for v in a.ravel()]}


def bip(b):
    return sorted(sorted(str(x) for x in side) for side in b)


def attrs(d):
    return {k: (fl(v) if isinstance(v, float) else v) for k, v in d.items()}


def tree_record(T, rng):
    n = T.size
    r = {"size": n}
    r["ancestors"] = [list(T.get_ancestors(i)) for i in range(n)]
    r["support"] = [fl(T.get_support(i)) for i in range(n)]
    internal = [int(i) for i in T.get_internal_nodes()]
    sub = internal[len(internal) // 2]  # some internal node (root for 3-node trees)
    r["sub"] = sub
    r["nodes"] = [int(i) for i in T.get_nodes()]
    r["internal"] = internal
    r["nodes_sub"] = [int(i) for i in T.get_nodes(sub)]
    r["internal_sub"] = [int(i) for i in T.get_internal_nodes(sub)]
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, size=(60, 2))]
    r["pairs"] = pairs
    r["is_descendant"] = [bool(T.is_descendant(a, b)) for a, b in pairs]
    r["is_ancestor"] = [int(T.is_ancestor(a, b)) for a, b in pairs]
    r["is_sibling"] = [bool(T.is_sibling(a, b)) for a, b in pairs]
    r["sibling_of_children"] = [bool(T.is_sibling(*T.get_children(i))) for i in internal]
    r["has_children"] = [bool(T.has_children(i)) for i in range(n)]
    r["has_parent"] = [bool(T.has_parent(i)) for i in range(n)]
    r["paths"] = [[int(x) for x in T.path_between_nodes(a, b)] for a, b in pairs]
    r["bipartitions_by_id"] = [bip(b) for b in T.bipartitions(by_id=True)]
    r["bipartitions"] = [bip(b) for b in T.bipartitions()]
    r["inorder"] = [[int(i), fl(d)] for i, d in T.traverse_inorder()]
    r["inorder_ids"] = [int(i) for i in T.traverse_inorder(include_distances=False)]
    for name in ("preorder", "postorder", "levelorder", "leaves_only", "internal_only"):
        fn = getattr(T, "traverse_" + name)
        r[name] = [int(i) for i in fn()]
        r[name + "_sub"] = [int(i) for i in fn(sub)]
    r["with_depth"] = [[int(i), int(d)] for i, d in T.traverse_with_depth()]
    r["with_depth_sub"] = [[int(i), int(d)] for i, d in T.traverse_with_depth(sub)]
    r["with_distances"] = [[int(i), fl(d), fl(c)] for i, d, c in T.traverse_with_distances()]
    r["with_distances_sub"] = [[int(i), fl(d), fl(c)] for i, d, c in T.traverse_with_distances(sub)]
    # (upstream shares one scratch buffer between these calls: after get_internal_nodes() it holds
    #  num_leaves slots and adjacency_matrix() overruns it; get_nodes() leaves it at full size)
    T.get_nodes()
    for tag, start in (("", None), ("_sub", sub)):
        A = T.adjacency_matrix(start)
        r["adjacency" + tag] = {"matrix": mat(A["adjacency_matrix"]), "node_ids": [int(i) for i in A["node_ids"]]}
        Lp = T.laplacian_matrix(start)
        r["laplacian" + tag] = {"matrix": mat(Lp["laplacian"]), "node_ids": [int(i) for i in Lp["node_ids"]]}
        D = T.degree_sequence(start)
        r["degrees" + tag] = {"degrees": [int(i) for i in D["degrees"]], "max": int(D["max_degree"]),
                              "min": int(D["min_degree"])}
    I = T.incidence_matrix()
    r["incidence"] = {"matrix": [int(v) for v in np.asarray(I["incidence_matrix"]).ravel()],
                      "shape": list(I["incidence_matrix"].shape), "node_ids": [int(i) for i in I["node_ids"]],
                      "edge_list": [[int(a), int(b)] for a, b in I["edge_list"]]}
    try:
        T.incidence_matrix(sub)
        r["incidence_sub_error"] = None
    except Exception as e:
        r["incidence_sub_error"] = type(e).__name__
    r["nx_nodes"] = [[int(i), attrs(a)] for i, a in T.to_networkx_nodes()]
    r["nx_edges"] = [[int(c), int(p), attrs(a)] for c, p, a in T.to_networkx_edges()]
    r["nx_nodes_sub"] = [[int(i), attrs(a)] for i, a in T.to_networkx_nodes(sub)]
    r["newick"] = T.to_newick()
    r["newick_plain"] = T.to_newick(include_support=False, include_distances=False)
    r["newick_sub"] = T.to_newick(sub)
    errs = {}
    leaf = int(T.leaf_node_ids[0])
    for what, fn in (("bipartition_of_leaf", lambda: T.bipartition(leaf)),
                     ("ancestors_out_of_range", lambda: list(T.get_ancestors(n))),
                     ("support_unknown_name", lambda: T.get_support("no such leaf")),
                     ("path_bad_type", lambda: T.path_between_nodes(1.5, 0))):
        try:
            fn()
            errs[what] = None
        except Exception as e:
            errs[what] = [type(e).__name__, str(e)]
    r["errors"] = errs
    return r


def main():
    import pandas as pd

    M = ref_loader.load_reference()
    rng = np.random.default_rng(11)
    with open(os.path.join(HERE, "trees.json")) as f:
        trees = json.load(f)
    out = {"trees": {}, "linked": {}}
    for name, rec in trees.items():
        src = rec.get("newick") or os.path.join(DATA, name)
        T = M.SuchTree(src)
        if T.size > 60:
            continue
        out["trees"][name] = tree_record(T, rng)
        print(name, T.size)
    for name, (t1, t2, lk) in {"gopher_louse": ("test.tree", "lice.tree", "links.csv"),
                               "fishworm": ("fishworm_host.tree", "fishworm_guest.tree", "fishworm_links.csv")}.items():
        T1, T2 = M.SuchTree(os.path.join(DATA, t1)), M.SuchTree(os.path.join(DATA, t2))
        links = pd.read_csv(os.path.join(DATA, lk), index_col=0)
        if set(links.index) != set(T1.leaves.keys()):
            links = links.T
        SLT = M.SuchLinkedTrees(T1, T2, links)
        rec = {}
        if name == "gopher_louse":
            rec["adjacency"] = mat(SLT.adjacency())
            rec["laplacian"] = mat(SLT.laplacian())
        else:  # 422 x 422: checksums only
            aj = SLT.adjacency()
            rec["adjacency_shape"] = list(aj.shape)
            rec["adjacency_row_sums"] = [fl(v) for v in aj.sum(axis=1)]
            rec["adjacency_nonzeros"] = int(np.count_nonzero(aj))
        rec["spectrum"] = [fl(v) for v in SLT.spectrum()]
        # a subset: the clade below an internal node of each tree
        sa = int(T1.get_internal_nodes()[2])
        sb = int(T2.get_internal_nodes()[1])
        SLT.subset_a(sa)
        SLT.subset_b(sb)
        T1.get_nodes()  # (restores upstream's shared scratch buffer to full size, see tree_record)
        T2.get_nodes()
        rec["subset"] = {"a": sa, "b": sb, "n_links": int(SLT.subset_n_links)}
        if SLT.subset_n_links > 0:
            aj = SLT.adjacency()
            rec["subset"]["adjacency_shape"] = list(aj.shape)
            rec["subset"]["adjacency_row_sums"] = [fl(v) for v in aj.sum(axis=1)]
            rec["subset"]["spectrum"] = [fl(v) for v in SLT.spectrum()]
        out["linked"][name] = rec
        print(name, rec.get("adjacency_shape") or rec["adjacency"]["shape"], rec["subset"])
    with open(os.path.join(HERE, "api.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
