"""CPU-only tests of the host side: the array-based NEWICK flattener against the
reference's structures, the synthetic generators, the C-ABI library (loads,
exports every declared symbol, fails loudly without a device), helpers."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from conftest import GOLDEN, REPO, load_pairs, read_tree_text, tree_source

import oracle as O
from suchtree_b200 import philox_host as philox_ref
import tree_build
from suchtree_b200 import _lib, newick, synth
from suchtree_b200.exceptions import InvalidNodeError, NodeNotFoundError, SuchTreeError, TreeStructureError
from suchtree_b200.tree import _read_tree_input


# ---------------------------------------------------------------- flattener --
def test_flattener_matches_reference(golden_trees):
    for name, rec in golden_trees.items():
        ft = newick.flatten(_read_tree_input(tree_source(name, rec)))
        assert ft.size == rec["size"] and ft.root == rec["root"] and ft.n_leaves == rec["num_leaves"], name
        assert list(ft.leaves.items()) == sorted(rec["leaves"].items(), key=lambda kv: kv[1]), name
        assert ft.parent.tolist() == rec["parent"], name
        assert ft.left.tolist() == rec["left"] and ft.right.tolist() == rec["right"], name
        assert [float(x).hex() for x in ft.distance] == rec["distance_hex"], name
        assert np.allclose(ft.support, rec["support"]), name


@pytest.mark.parametrize("name", ["ml", "nj"])
def test_flattener_big_trees(name, bigtrees):
    ft = newick.flatten(read_tree_text("%s.tree.gz" % name))
    z = np.load(os.path.join(GOLDEN, "pairs_big_%s.npz" % name))
    info = bigtrees[name]
    assert (ft.size, ft.root, ft.n_leaves) == (info["size"], info["root"], info["num_leaves"])
    assert np.array_equal(ft.parent, z["parent"])
    for k, v in list(info["first_leaves"].items()) + list(info["last_leaves"].items()):
        assert ft.leaves[k] == v


def test_flattener_agrees_with_oracle_builder_on_odd_inputs():
    cases = [
        "A;",
        "(A,B);",
        "(A:1,(B:2,C:3):4,(D:5,E:6,F:7,G:8):9,H:10)r:11;",
        "((((a,b),c),d),(e,(f,(g,(h,i)))));",
        " ( 'x y' : 1.5 , ( z:2 , w : 3 ) 0.75 : 4 ) ; ",
        "(A:1,B:2)[comment](C:3);" if False else "((A:1,B:2)[c]:1,(C:3,D:4):2);",
        "(a:1e-3,b:1E2,(c:-1,d:+2):0.0);",
    ]
    for nw in cases:
        ft = newick.flatten(nw)
        a = tree_build.build_arrays(nw.strip()) if nw.strip().startswith("(") else None
        if a is None:
            assert ft.size == 1 and ft.root == 0 and ft.leaves == {"A": 0}
            continue
        assert ft.parent.tolist() == a["parent"].tolist(), nw
        assert ft.left.tolist() == a["left"].tolist(), nw
        assert ft.leaves == a["leaves"], nw
        assert [float(x).hex() for x in ft.distance] == [float(x).hex() for x in a["distance"]], nw


def test_flattener_rejects_bad_trees():
    for bad in ["((A,B));", "(A,(B));", "((A,B),(C,D)", "(A,B));", "", "(A,,B);"]:
        with pytest.raises(TreeStructureError):
            newick.flatten(bad)


def _same_flat(a, b):
    return (a.size == b.size and a.root == b.root and a.n_leaves == b.n_leaves
            and np.array_equal(a.parent, b.parent) and np.array_equal(a.left, b.left)
            and np.array_equal(a.right, b.right)
            and np.array_equal(a.distance.view(np.uint32), b.distance.view(np.uint32))
            and np.array_equal(a.support.view(np.uint32), b.support.view(np.uint32))
            and list(a.leaves.items()) == list(b.leaves.items())
            and np.array_equal(a.internal_nodes, b.internal_nodes))


def test_native_loader_equals_python_restatement(golden_trees):
    """csrc/st_newick.cu (what the product runs) against newick.flatten_py (the same
    rules in Python), field for field and bit for bit."""
    texts = [_read_tree_input(tree_source(name, rec)) for name, rec in golden_trees.items()]
    texts += [read_tree_text("ml.tree.gz"), read_tree_text("nj.tree.gz")]
    texts += [
        "A;", "(A,B);", "('':1,B:2);", "(A:1,B:2,C:3,D:4,E:5,F:6,G:7)0.5:3;",
        "((A:1,B:2)90:1,(C:3,D:4)'x':2)root;", "(A:inf,B:-0.0,(C:1e400,D:NaN):1e-400);",
        " ( 'x y' : 1.5 , ( z:2 , w : 3 ) 0.75 : 4 ) ; trailing (junk", "[c](A[x]:1[y],B:2)[z];",
        "(A:1,B:2);(C,D);", "((a,b,c,d,e,f,g,h,i,j)1e2:1,(k,l,m)x1:2,n);", "(\u00e9t\u00e9:1,'\u65e5\u672c':2);",
        "(A:.5,B:5.,(C:+1.e+1,D:-.5E-1):1);",
        "(A:1,(Cuphea_o'donellii:1,B''s:2):1,'it''s':3);",
    ]
    for t in texts:
        assert _same_flat(newick.flatten(t), newick.flatten_py(t)), t[:60]
    rng = np.random.default_rng(3)
    for n in (2, 3, 17, 400):
        ft0 = synth.yule_tree(n, seed=int(rng.integers(1 << 30)), names=True)
        nw = synth.to_newick(ft0)
        a = newick.flatten(nw)
        assert _same_flat(a, newick.flatten_py(nw)) and np.array_equal(a.parent, ft0.parent)


def test_quote_inside_a_bare_label_is_an_ordinary_character():
    """data/plant-pollinators/rabr/plant.tree of the reference has the leaf
    Cuphea_o'donellii, and its link table (made with real dendropy) uses that name: a
    quote opens a quoted label only at the start of a token."""
    import tree_build

    nw = "((Ludwigia_nervosa:1,Cuphea_o'donellii:1)1:1,('quoted name':2,x'y'z:1):1);"
    for ft in (newick.flatten(nw), newick.flatten_py(nw)):
        assert list(ft.leaves) == ["Ludwigia_nervosa", "Cuphea_o'donellii", "quoted name", "x'y'z"]
    assert tree_build.build_arrays(nw)["leaves"] == newick.flatten(nw).leaves


def test_native_loader_error_messages():
    for bad in ["((A,B));", "(A,(B));", "((A,B),(C,D)", "(A,B));", "", "(A,,B);", "(A:x,B:1);", "(A:'1',B:1);",
                "[only a comment]", "(A,B),C;"]:
        with pytest.raises(TreeStructureError) as e1:
            newick.flatten(bad)
        with pytest.raises(TreeStructureError) as e2:
            newick.flatten_py(bad)
        assert str(e1.value)[:25] == str(e2.value)[:25], bad


def test_deep_caterpillar_newick_does_not_recurse():
    L = 20000
    ft0 = synth.caterpillar_tree(L, seed=3, names=True)
    ft1 = newick.flatten(synth.to_newick(ft0))
    assert np.array_equal(ft0.parent, ft1.parent) and np.array_equal(ft0.left, ft1.left)
    assert np.array_equal(ft0.distance[ft0.parent >= 0], ft1.distance[ft1.parent >= 0])


# --------------------------------------------------------------- generators --
def _check_inorder(ft):
    """ids are in-order ranks <=> every left subtree precedes its node, right follows."""
    n = ft.size
    assert (ft.parent == -1).sum() == 1 and ft.parent[ft.root] == -1
    internal = np.nonzero(ft.left != -1)[0]
    assert np.all(internal % 2 == 1) and np.all(np.nonzero(ft.left == -1)[0] % 2 == 0)
    assert np.all(ft.parent[ft.left[internal]] == internal) and np.all(ft.parent[ft.right[internal]] == internal)
    assert np.all(ft.left[internal] < internal) and np.all(ft.right[internal] > internal)
    # depth-argmin property on random pairs against the parent-walking oracle
    t = O.OracleTree(ft.parent, ft.distance)
    depth = t.node_depths()
    rng = np.random.default_rng(1)
    pairs = rng.integers(0, n, size=(300, 2))
    m = (t.distances_f64_climb(pairs, with_mrca=True))[1]
    for (a, b), mm in zip(pairs, m):
        lo, hi = min(a, b), max(a, b)
        assert lo + int(np.argmin(depth[lo:hi + 1])) == mm


@pytest.mark.parametrize("gen", [synth.yule_tree, synth.balanced_tree, synth.caterpillar_tree])
@pytest.mark.parametrize("n_leaves", [1, 2, 3, 7, 64, 1000, 4097])
def test_generators_produce_valid_inorder_trees(gen, n_leaves):
    ft = gen(n_leaves, seed=5)
    assert ft.size == 2 * n_leaves - 1
    if n_leaves == 1:
        assert ft.root == 0 and ft.left[0] == -1
        return
    _check_inorder(ft)
    e = ft.distance[ft.parent >= 0]
    assert e.dtype == np.float32 and np.all(e >= 0.5) and np.all(e < 1.0)
    assert np.array_equal(gen(n_leaves, seed=5).parent, ft.parent)  # deterministic


def test_generator_shapes():
    assert O.OracleTree(*_pd(synth.balanced_tree(1 << 12))).depth == 13
    assert O.OracleTree(*_pd(synth.caterpillar_tree(500))).depth == 500
    d = O.OracleTree(*_pd(synth.yule_tree(100000, seed=1))).depth
    assert 25 <= d <= 80  # ~2 ln L expected mean leaf depth


def _pd(ft):
    return ft.parent, ft.distance


def test_newick_roundtrip_through_oracle_builder():
    ft = synth.yule_tree(300, seed=9, names=True)
    a = tree_build.build_arrays(synth.to_newick(ft))
    assert np.array_equal(a["parent"], ft.parent) and a["leaves"] == ft.leaves
    assert np.array_equal(a["distance"][a["parent"] >= 0], ft.distance[ft.parent >= 0])


# ------------------------------------------------------------------- C ABI ---
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "suchtree_b200.h")).read()
    declared = set(re.findall(r"^ST_API [^;(]*?\b(st_[a-z0-9_]+)\(", header, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.st_version() >= 100
    # the binary was compiled from the sources on disk (a stale .so is rebuilt, never used)
    from suchtree_b200 import build as B

    assert L.st_build_id().decode() == B.source_id("product")


def test_bench_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "suchtree_b200_bench.h")).read()
    declared = set(re.findall(r"^ST_BENCH_API [^;(]*?\b(st_[a-z0-9_]+)\(", header, flags=re.M))
    assert declared == set(_lib.BENCH_SIGNATURES), declared ^ set(_lib.BENCH_SIGNATURES)
    B = _lib.bench_lib()
    for name in declared:
        assert hasattr(B, name), name
    # measurement tooling stays out of the product library
    assert not hasattr(_lib.lib(), "st_bench_gather")


def test_no_cpu_fallback_without_device():
    n = C.c_int(0)
    rc = _lib.lib().st_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    from suchtree_b200 import SuchTree, pearson

    with pytest.raises(SuchTreeError):
        SuchTree("(A,B,(C,D));")
    with pytest.raises(SuchTreeError):
        pearson(np.arange(4.0), np.arange(4.0))


def test_tree_validation_happens_before_any_device_work():
    L = _lib.lib()
    h = C.c_void_p()

    def create(parent, left, right):
        p, l, r = (np.array(x, np.int32) for x in (parent, left, right))
        e = np.ones(len(parent), np.float32)
        return L.st_tree_create(0, len(parent), p.ctypes.data, l.ctypes.data, r.ctypes.data, e.ctypes.data, 0, 0, C.byref(h))

    # not in-order: root stored first
    assert create([-1, 0, 0], [1, -1, -1], [2, -1, -1]) == _lib.ST_ERR_NOT_INORDER
    # unary node
    assert create([1, -1], [-1, 0], [-1, -1]) == _lib.ST_ERR_NOT_BINARY
    # two roots
    assert create([-1, -1, 1], [-1, 0, -1], [-1, 2, -1]) in (_lib.ST_ERR_NOT_BINARY, _lib.ST_ERR_INVALID_ARG)
    assert b"" != L.st_last_error()


def test_exceptions_mirror_reference():
    e = InvalidNodeError(7, 5)
    assert str(e) == "Node ID 7 out of bounds (tree size: 5)" and e.node_id == 7 and e.tree_size == 5
    assert str(InvalidNodeError(3)) == "Invalid node ID: 3"
    assert str(NodeNotFoundError("x")) == "Leaf name not found: x."
    assert str(NodeNotFoundError(4)) == "Node not found: 4"
    assert issubclass(TreeStructureError, SuchTreeError) and issubclass(SuchTreeError, Exception)


# ------------------------------------------------------------------ Philox ---
def test_philox_known_answers():
    """Random123 kat_vectors: philox4x32-10, zero counter/key and the pi-digits vector."""
    x = philox_ref.philox4x32_10(np.array([0], np.uint64), 0)
    assert [int(v[0]) for v in x] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    # counter words (243f6a88, 85a308d3, 13198a2e, 03707344) need c2,c3 != 0: use the
    # generic round function directly
    c = [np.array([v], np.uint64) for v in (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)]
    k0, k1 = 0xA4093822, 0x299F31D0
    for _ in range(10):
        p0 = philox_ref.M0 * c[0]
        p1 = philox_ref.M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & philox_ref.MASK,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & philox_ref.MASK]
        k0 = (k0 + philox_ref.W0) & 0xFFFFFFFF
        k1 = (k1 + philox_ref.W1) & 0xFFFFFFFF
    assert [int(v[0]) for v in c] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_random_leaf_pairs_reference_is_shardable():
    a = philox_ref.random_leaf_pairs(1000, 42, 0, 101)
    b = np.concatenate([philox_ref.random_leaf_pairs(1000, 42, 0, 40), philox_ref.random_leaf_pairs(1000, 42, 40, 61)])
    assert np.array_equal(a, b)
    assert a.min() >= 0 and a.max() <= 1998 and np.all(a % 2 == 0)


def _random_newick(rng, n_leaves):
    """NEWICK text with the syntax the reference's inputs show: polytomies, missing / zero /
    negative / scientific lengths, quoted labels (with '' escapes), a quote inside a bare label,
    comments, support values as internal labels, stray blanks."""
    count = [0]

    def label():
        count[0] += 1
        i, k = count[0], rng.random()
        if k < 0.6:
            return "t%d" % i
        if k < 0.7:
            return "sp_%d_x" % i
        if k < 0.8:
            return "'quoted %d'" % i
        if k < 0.85:
            return "'it''s %d'" % i
        if k < 0.9:
            return "a'b%d" % i
        return "T%d.1" % i

    def length():
        k = rng.random()
        if k < 0.15:
            return ""
        if k < 0.3:
            return ":0" if k < 0.25 else ":0.0"
        if k < 0.4:
            return ":%.3e" % rng.uniform(1e-6, 10)
        if k < 0.45:
            return ":-%.4f" % rng.uniform(0, 1)
        return ":%.6f" % rng.uniform(0, 2)

    def support():
        k = rng.random()
        if k < 0.5:
            return ""
        if k < 0.7:
            return "%d" % rng.randint(0, 100)
        if k < 0.85:
            return "%.3f" % rng.random()
        return "[&c=1]" if k < 0.9 else "n%d" % rng.randint(0, 9)

    def node(budget):
        if budget == 1:
            return label() + ("[comment]" if rng.random() < 0.1 else "") + length()
        k = 2 if rng.random() < 0.7 or budget < 3 else rng.randint(3, min(6, budget))
        cuts = sorted(rng.sample(range(1, budget), k - 1))
        parts = [b - a for a, b in zip([0] + cuts, cuts + [budget])]
        sep = ", " if rng.random() < 0.1 else ","
        return "(" + sep.join(node(p) for p in parts) + ")" + support() + length()

    return node(n_leaves) + ";"


def test_three_loaders_agree_on_random_newick():
    """Native loader (csrc/st_newick.cu), its Python restatement and the oracle's builder (the
    dendropy stand-in + the restated __init__ passes, pinned to the reference's structures by
    test_oracle_golden.py) on 200 seeded random trees: ids, structure, fp32 lengths bit for bit,
    leaf names, support values.  (The same generator, run against the compiled reference in
    the authoring container over 1,100 trees, found no difference.)"""
    import random

    import tree_build

    rng = random.Random(20261017)
    for _ in range(200):
        text = _random_newick(rng, rng.randint(2, 40))
        a = tree_build.build_arrays(text)
        for ft in (newick.flatten(text), newick.flatten_py(text)):
            assert int(ft.size) == a["size"] and int(ft.root) == a["root"], text
            assert dict(ft.leaves) == a["leaves"], text
            for k in ("parent", "left", "right"):
                assert np.array_equal(getattr(ft, k), a[k]), (k, text)
            assert np.asarray(ft.distance, np.float32).tobytes() == np.asarray(a["distance"], np.float32).tobytes(), text
        assert np.array_equal(newick.flatten(text).support, newick.flatten_py(text).support, equal_nan=True), text


def test_input_registration_bookkeeping_without_a_device(monkeypatch):
    """maybe_register(): thresholds, the policy variable, owners it cannot track, and the
    weakref finaliser that forgets an array when it dies (no device here: the registration
    call itself fails and is not retried on every call)."""
    import gc

    monkeypatch.setattr(_lib, "REGISTER_MIN_BYTES", 1 << 16)
    calls = []

    class Fake:
        def st_host_register(self, p, n):
            calls.append(("reg", p, n))
            return 0

        def st_host_unregister(self, p):
            calls.append(("unreg", p))
            return 0

    monkeypatch.setattr(_lib, "lib", lambda: Fake())
    small = np.zeros((100, 2), np.int64)
    assert _lib.maybe_register(small) is False and not calls
    big = np.zeros((1 << 13, 2), np.int64)  # 128 KiB
    monkeypatch.setenv("SUCHTREE_B200_REGISTER", "0")
    assert _lib.maybe_register(big) is False and not calls
    monkeypatch.setenv("SUCHTREE_B200_REGISTER", "2")
    assert _lib.maybe_register(big) is False and not calls  # first sighting
    assert _lib.maybe_register(big[: 1 << 12]) is True  # second sighting (through a view): the OWNER is registered
    assert calls == [("reg", big.ctypes.data, big.nbytes)]
    assert _lib.maybe_register(big) is True and len(calls) == 1
    # memory owned by something that is not an ndarray is left alone
    raw = bytearray(1 << 18)
    foreign = np.frombuffer(raw, dtype=np.int64).reshape(-1, 2)
    assert _lib.maybe_register(foreign) is False and _lib.maybe_register(foreign) is False and len(calls) == 1
    key, ptr = id(big), big.ctypes.data
    assert key in _lib._registered
    del big
    gc.collect()
    assert key not in _lib._registered and calls[-1] == ("unreg", ptr)


# ------------------------------------------------------------- name -> id glue (C) ----------
def test_py_glue_converts_name_rows_and_reports_the_first_bad_row():
    """pyglue/st_pynames.c: the by-name entry points' name -> id walk as one C loop
    (MuchTree.pyx:964-975, :1397-1408).  -1 = all rows converted; otherwise the index of the
    first row the reference's own loop has to look at (it raises the reference's error)."""
    import ctypes as C

    from suchtree_b200 import _lib

    g = _lib.py_glue()
    if g is None:
        pytest.skip("no C compiler / Python.h here: the shim keeps its Python walk")
    leaves = {"n%d" % i: 2 * i for i in range(1000)}
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 1000, size=(5000, 2))
    rows = [("n%d" % a, "n%d" % b) for a, b in idx]
    out = np.full((5000, 2), -7, dtype=np.int64)
    assert g.st_py_names_to_ids(rows, leaves, out.ctypes.data, 2) == -1
    assert np.array_equal(out, 2 * idx)
    rows4 = [["n1", "n2", "n3", "n4"], ("n5", "n6", "n7", "n8")]  # lists and tuples both count as rows
    out4 = np.empty((2, 4), dtype=np.int64)
    assert g.st_py_names_to_ids(rows4, leaves, out4.ctypes.data, 4) == -1
    assert out4.tolist() == [[2, 4, 6, 8], [10, 12, 14, 16]]
    for bad_row, at in ((("n1", "nope"), 17), (("n1", 3), 4), (("n1",), 0), ("n1n2", 9), (("n1", "n2", "n3"), 4999)):
        r = list(rows)
        r[at] = bad_row
        assert g.st_py_names_to_ids(r, leaves, out.ctypes.data, 2) == at
    assert g.st_py_names_to_ids([], leaves, out.ctypes.data, 2) == -1
    assert g.st_py_names_to_ids(tuple(rows), leaves, out.ctypes.data, 2) == 0  # not a list: caller's slow path

    class Name(str):
        pass

    assert g.st_py_names_to_ids([(Name("n1"), "n2")], leaves, out.ctypes.data, 2) == 0  # exact str only


# ------------------------------------------------------------- bit-packed pair stream -------
@pytest.mark.parametrize("w", [1, 2, 7, 17, 18, 21, 31])
def test_host_packer_writes_the_bit_stream_the_pair_kernel_reads(w):
    """st_host_pack_pairs (the host stage of st_distances, MuchTree.pyx:872-909 takes int64 pairs):
    pair i occupies bits [i*2w, (i+1)*2w), a low, b high; any n (word boundaries, parts of the
    thread pool), any strides; the OR of the ids flags what cannot be a node id."""
    import ctypes as C

    from suchtree_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(w)
    for n in (1, 2, 63, 64, 65, 1000, 32768 * 3 + 17, 400_001):
        ids = rng.integers(0, 1 << w, size=(n, 2), dtype=np.int64)
        words = (n * 2 * w + 63) // 64 + 1
        for arr in (ids, np.asfortranarray(ids), np.ascontiguousarray(ids[:, ::-1])[:, ::-1]):
            out = np.full(words, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
            acc = C.c_uint64()
            rc = L.st_host_pack_pairs(arr.ctypes.data, arr.strides[0] // 8, arr.strides[1] // 8, n, w,
                                      out.ctypes.data, C.byref(acc))
            assert rc == 0 and acc.value >> w == 0
            bits = np.unpackbits(out[:-1].view(np.uint8), bitorder="little")[: n * 2 * w].reshape(n, 2, w)
            got = (bits.astype(np.int64) << np.arange(w, dtype=np.int64)).sum(axis=2)
            assert np.array_equal(got, ids), (w, n)
    bad = np.array([[0, 1], [1 << w, 0]], dtype=np.int64)
    out = np.zeros(4, dtype=np.uint64)
    acc = C.c_uint64()
    assert L.st_host_pack_pairs(bad.ctypes.data, 2, 1, 2, w, out.ctypes.data, C.byref(acc)) == 0 and acc.value >> w
    bad[1, 0] = -1
    assert L.st_host_pack_pairs(bad.ctypes.data, 2, 1, 2, w, out.ctypes.data, C.byref(acc)) == 0 and acc.value >> w
