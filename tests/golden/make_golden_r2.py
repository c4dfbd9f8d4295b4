"""Golden vectors added in round 2, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_r2.py  ->  tests/golden/r2.json

per small tree: list(get_descendants(node)) for every node (the reference's generator order,
MuchTree.pyx:396-425), and distance_to_root of every node of a tree that has a real edge of
length exactly -1 (the walk of MuchTree.pyx:845-849 stops there).
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
MINUS_ONE = "((A:1,B:-1):0.5,(C:0.25,(D:2,E:-1):-1):1.5);"


def main():
    M = ref_loader.load_reference()
    with open(os.path.join(HERE, "trees.json")) as f:
        trees = json.load(f)
    out = {"descendants": {}, "minus_one": {"newick": MINUS_ONE}}
    for name, rec in trees.items():
        src = rec.get("newick") or os.path.join(DATA, name)
        T = M.SuchTree(src)
        if T.size > 80:
            continue
        out["descendants"][name] = [[int(v) for v in T.get_descendants(i)] for i in range(T.size)]
    T = M.SuchTree(MINUS_ONE)
    out["minus_one"]["distance_to_root"] = [float(T.distance_to_root(i)) for i in range(T.size)]
    with open(os.path.join(HERE, "r2.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out["descendants"].items()}, out["minus_one"])


if __name__ == "__main__":
    main()
