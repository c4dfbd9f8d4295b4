"""SASS evidence for the hot kernels (runs here: cuobjdump needs no GPU):

    python scripts/sass_excerpts.py > profiles/r02_sass_excerpts.txt

Per kernel: instruction count, the mnemonics that show how memory is touched (UBLKCP = the
TMA bulk copy that stages the block tables / matrix column blocks, SYNCS = its mbarrier,
LDG.E...128/256 = one-sector record gathers and streaming id loads, STG...  = streaming
result stores with their cache hints), and the lines themselves."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "suchtree_b200", "libsuchtree_b200.so")
WANT = [r"k_pairsIiLi2ELi1ELi512ELi2ELi0E", r"k_pairsIiLi2ELi1ELi512ELi2ELi1E", r"k_pairsIlLi2ELi1ELi512ELi2ELi0E",
        r"k_pairsIiLi2ELi0ELi512ELi2ELi0E", r"k_matrix_ordered", r"k_matrix_diag", r"k_quartetsILi1ElLi1ELi3ELb1ELb0ElLi384E",
        r"k_quartetsILi1EiLi1ELi4ELb1ELb0EiLi256E", r"k_sample_momentsILi1ELi1E", r"k_linked_momentsILi1ELi1E",
        r"k_clade_momentsILi1ELi1E", r"k_sample_xs", r"k_bucket_sums"]
PAT = re.compile(r"\b(UBLKCP|UTMALDG|SYNCS|LDG\.E[\w.]*|STG\.E[\w.]*|LDS[\w.]*|ATOMG[\w.]*|RED[\w.]*|LDGSTS[\w.]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    print("# cuobjdump -sass suchtree_b200/libsuchtree_b200.so (sm_100a), excerpts per hot kernel")
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if not any(re.search(w, name) for w in WANT):
            continue
        lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        counts = {}
        shown = []
        for l in lines:
            m = PAT.search(l)
            if m:
                counts[m.group(1)] = counts.get(m.group(1), 0) + 1
                if not m.group(1).startswith("LDS"):
                    shown.append(re.sub(r"\s*/\*[0-9a-f]{16}\*/\s*$", "", l).rstrip())
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print("\n## %s\n   %s\n   %d SASS instructions; %s" % (
            name, dem, len(lines), ", ".join("%s x%d" % kv for kv in sorted(counts.items()))))
        for l in shown[:40]:
            print(l)
        if len(shown) > 40:
            print("        ... %d more" % (len(shown) - 40))


if __name__ == "__main__":
    main()
