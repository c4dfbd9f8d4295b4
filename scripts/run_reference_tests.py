"""The reference's OWN test-suite (SuchTree/tests/*.py, 117 tests) against suchtree_b200.

    python scripts/run_reference_tests.py stage     # here (needs /root/reference): stage the files
    gpurun -- 'python scripts/run_reference_tests.py run'
    python scripts/run_reference_tests.py clean     # remove the staged copy again

`stage` copies the reference's test files and their data UNMODIFIED into oracle/_ref/reftests/
(git-ignored: nothing of the reference enters the history; the directory travels to the GPU box
with the gpurun snapshot) and writes a conftest.py there that makes `import SuchTree` resolve
to suchtree_b200 (SuchTree, SuchLinkedTrees, the four exception classes) and `import dendropy`
to the stand-in the oracle uses (oracle/newick_ref.py; the reference's tests use dendropy only to
cross-check child ids).  `run` runs pytest on the staged copy.

Round 2, B200: 116 passed, 1 skipped (to_igraph: igraph is not installed) -- the same outcome as
the reference's tests against the reference itself.
"""
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(REPO, "oracle", "_ref", "reftests")
REF_TESTS = "/root/reference/SuchTree/tests"

CONFTEST = '''# written by scripts/run_reference_tests.py: `import SuchTree` -> suchtree_b200
import os, sys, types
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
import newick_ref
dp = types.ModuleType("dendropy")
dp.Tree, dp.Node, dp.Taxon = newick_ref.Tree, newick_ref.Node, newick_ref.Taxon
sys.modules["dendropy"] = dp
import suchtree_b200 as ours
from suchtree_b200 import exceptions as ex
pkg = types.ModuleType("SuchTree")
pkg.__path__ = []
for name in ("SuchTree", "SuchLinkedTrees"):
    setattr(pkg, name, getattr(ours, name))
for name in ("SuchTreeError", "NodeNotFoundError", "InvalidNodeError", "TreeStructureError"):
    setattr(pkg, name, getattr(ex, name))
pkg.exceptions = ex
sys.modules["SuchTree"] = pkg
sys.modules["SuchTree.exceptions"] = ex
'''


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "run"
    if what == "stage":
        shutil.rmtree(STAGE, ignore_errors=True)
        dst = os.path.join(STAGE, "SuchTree", "tests")  # the tests open 'SuchTree/tests/<file>' relative to the cwd
        os.makedirs(dst)
        for name in os.listdir(REF_TESTS):
            if os.path.isfile(os.path.join(REF_TESTS, name)):
                shutil.copy(os.path.join(REF_TESTS, name), dst)
        with open(os.path.join(STAGE, "conftest.py"), "w") as f:
            f.write(CONFTEST)
        print("staged", len(os.listdir(dst)), "files under", STAGE)
    elif what == "clean":
        shutil.rmtree(STAGE, ignore_errors=True)
    else:
        if not os.path.isdir(STAGE):
            sys.exit("nothing staged: run `python scripts/run_reference_tests.py stage` where /root/reference exists")
        sys.exit(subprocess.call([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-W",
                                  "ignore::DeprecationWarning", "SuchTree/tests"], cwd=STAGE))


if __name__ == "__main__":
    main()
