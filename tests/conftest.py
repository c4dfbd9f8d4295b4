import gzip
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
DATA = os.path.join(GOLDEN, "data")
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes as C

        from suchtree_b200 import _lib

        n = C.c_int(0)
        return _lib.lib().st_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: the product has no
    # CPU fallback.  Without -m, GPU tests are skipped when no device is visible.
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def read_tree_text(name):
    p = os.path.join(DATA, name)
    if name.endswith(".gz"):
        with gzip.open(p, "rt") as f:
            return f.read()
    with open(p) as f:
        return f.read()


@pytest.fixture(scope="session")
def golden_trees():
    with open(os.path.join(GOLDEN, "trees.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def bigtrees():
    with open(os.path.join(GOLDEN, "bigtrees.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def published():
    with open(os.path.join(GOLDEN, "published.json")) as f:
        return json.load(f)


def tree_source(name, rec):
    """What the reference's constructor was given: the inline NEWICK string, or the
    path of the fixture file."""
    return rec.get("newick") or os.path.join(DATA, name)


def load_pairs(name):
    return np.load(os.path.join(GOLDEN, "pairs_%s.npz" % name))


def matrix_rows():
    rows = []
    with open(os.path.join(DATA, "test.matrix")) as f:
        for line in f:
            a, b, d = line.split()
            rows.append((a, b, float(d)))
    return rows
