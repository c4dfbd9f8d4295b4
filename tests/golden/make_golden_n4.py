"""Golden vectors for the thin callers of §8f N4, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_n4.py  ->  tests/golden/n4.json

per small tree: distance_to_root of every node, the RED dictionary (as [node, value]
pairs in insertion order), nearest_neighbors of every leaf (k=3), and relationships()
rows keyed by the unordered leaf pair.
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)
import ref_loader  # noqa: E402

warnings.simplefilter("ignore", DeprecationWarning)
DATA = os.path.join(HERE, "data")


def main():
    M = ref_loader.load_reference()
    with open(os.path.join(HERE, "trees.json")) as f:
        trees = json.load(f)
    out = {}
    for name, rec in trees.items():
        src = rec.get("newick") or os.path.join(DATA, name)
        T = M.SuchTree(src)
        if T.size > 80:
            continue
        r = {"distance_to_root": [float(T.distance_to_root(i)) for i in range(T.size)]}
        try:
            red = T.relative_evolutionary_divergence
            r["red"] = [[int(k), float(v)] for k, v in red.items()]
        except Exception as e:  # a+b == 0 somewhere (zero-length cherries): the reference raises
            r["red_error"] = str(e)[:40]
        r["nearest"] = {leaf: [[nm, float(d)] for nm, d in T.nearest_neighbors(leaf, k=3)] for leaf in T.leaves}
        try:
            df = T.relationships()
            r["relationships"] = {
                "|".join(sorted((row.a, row.b))): [row.a, row.b, float(row.distance), float(row.a_to_root),
                                                   float(row.b_to_root), int(row.mrca), float(row.mrca_to_root),
                                                   float(row.a_to_mrca), float(row.b_to_mrca)]
                for row in df.itertuples()}
        except Exception as e:
            r["relationships_error"] = repr(e)[:80]
        out[name] = r
        print(name, T.size, [k for k in r])
    with open(os.path.join(HERE, "n4.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
