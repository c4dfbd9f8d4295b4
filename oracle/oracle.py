"""TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/st_oracle.c (the CPU restatement of the reference's
hot path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package
(suchtree_b200/) never does and fails loudly without its CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "st_oracle.c")
OUT = os.path.join(HERE, "_build", "libst_oracle.so")
_lib = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"])
    return OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_mrca.restype = C.c_int
        L.oracle_mrca.argtypes = [i32p, i64p, C.c_int, C.c_int]
        L.oracle_mrca_bulk.argtypes = [i32p, i64p, i64p, C.c_uint64, i32p]
        L.oracle_distances_f32.argtypes = [i32p, f32p, i64p, i64p, C.c_uint64, f64p]
        L.oracle_distances_f64.argtypes = [i32p, f32p, i64p, i64p, C.c_uint64, f64p, C.c_void_p]
        L.oracle_distances_f64_climb.argtypes = [i32p, i32p, f32p, i64p, C.c_uint64, f64p, C.c_void_p]
        L.oracle_quartet_topologies.argtypes = [i32p, i64p, i64p, C.c_uint64, i64p]
        L.oracle_tree_depth.restype = C.c_uint
        L.oracle_tree_depth.argtypes = [i32p, i64p, C.c_uint64]
        L.oracle_node_depths.argtypes = [i32p, C.c_uint64, i32p]
        L.oracle_pearson_f32.restype = C.c_double
        L.oracle_pearson_f32.argtypes = [f64p, f64p, C.c_uint]
        L.oracle_pearson_f64.restype = C.c_double
        L.oracle_pearson_f64.argtypes = [f64p, f64p, C.c_uint64]
        L.oracle_linked_pairs.argtypes = [i64p, C.c_uint64, i64p, i64p]
        L.oracle_random_int.restype = C.c_uint64
        L.oracle_random_int.argtypes = [C.POINTER(C.c_uint64), C.c_uint64]
        L.oracle_sample_bucket.argtypes = [C.POINTER(C.c_uint64), i64p, C.c_uint64, C.c_uint, i64p, i64p]
        _lib = L
    return _lib


def _pairs(ids):
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    assert ids.ndim == 2 and ids.shape[1] == 2
    return ids


class OracleTree:
    """The reference's Node array, as flat arrays, plus its cdef kernels."""

    def __init__(self, parent, distance, leaf_ids=None, depth=None):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.distance = np.ascontiguousarray(distance, dtype=np.float32)
        self.size = int(self.parent.shape[0])
        self._node_depth = None
        if depth is None:
            # max #nodes on a leaf->root path.  Computed by exact integer pointer
            # jumping: the literal walk (literal_depth below, MuchTree.pyx:218-225) is
            # O(leaves x depth) and cannot finish on a 10^6-deep caterpillar.
            depth = int(self.node_depths().max()) + 1
        self.depth = int(depth)

    def literal_depth(self, leaf_ids=None):
        """MuchTree.pyx:218-225 restated literally (small trees only)."""
        if leaf_ids is None:
            has_child = np.zeros(self.size, bool)
            has_child[self.parent[self.parent >= 0]] = True
            leaf_ids = np.nonzero(~has_child)[0]
        leaf_ids = np.ascontiguousarray(leaf_ids, dtype=np.int64)
        return int(lib().oracle_tree_depth(self.parent, leaf_ids, leaf_ids.shape[0]))

    @classmethod
    def from_newick(cls, tree_input):
        import tree_build

        a = tree_build.build_arrays(tree_input)
        t = cls(a["parent"], a["distance"], depth=a["depth"])
        t.arrays = a
        return t

    def _visited(self):
        return np.zeros(max(self.depth, 1), np.int64)  # MuchTree.pyx:906

    def mrca(self, a, b):
        return int(lib().oracle_mrca(self.parent, self._visited(), int(a), int(b)))

    def mrca_bulk(self, ids):
        ids = _pairs(ids)
        out = np.empty(ids.shape[0], np.int32)
        lib().oracle_mrca_bulk(self.parent, self._visited(), ids, ids.shape[0], out)
        return out

    def quartet_topologies(self, quartets):
        """MuchTree.pyx:1331-1376."""
        q = np.ascontiguousarray(quartets, dtype=np.int64)
        assert q.ndim == 2 and q.shape[1] == 4
        out = np.zeros_like(q)
        lib().oracle_quartet_topologies(self.parent, self._visited(), q, q.shape[0], out)
        return out

    def distances_f32(self, ids):
        """O1': the reference's literal fp32-accumulating arithmetic."""
        ids = _pairs(ids)
        out = np.zeros(ids.shape[0], np.float64)
        lib().oracle_distances_f32(self.parent, self.distance, self._visited(), ids, ids.shape[0], out)
        return out

    def distances_f64(self, ids, with_l1=False):
        """O2: same path summation, fp64 accumulator over fp32-quantised edges."""
        ids = _pairs(ids)
        out = np.zeros(ids.shape[0], np.float64)
        l1 = np.zeros(ids.shape[0], np.float64) if with_l1 else None
        lib().oracle_distances_f64(
            self.parent, self.distance, self._visited(), ids, ids.shape[0], out,
            l1.ctypes.data if with_l1 else None,
        )
        return (out, l1) if with_l1 else out

    def node_depths(self):
        if self._node_depth is None:
            # vectorised pointer-jumping equivalent of oracle_node_depths (exact, ints)
            par = self.parent.astype(np.int64)
            depth = (par >= 0).astype(np.int64)
            anc = par.copy()
            while True:
                live = anc >= 0
                if not live.any():
                    break
                depth[live] += depth[anc[live]]
                anc[live] = anc[anc[live]]
            self._node_depth = depth.astype(np.int32)
        return self._node_depth

    def distances_f64_climb(self, ids, with_mrca=False):
        """O(depth) variant for trees the reference's O(depth^2) scan cannot finish."""
        ids = _pairs(ids)
        out = np.zeros(ids.shape[0], np.float64)
        m = np.zeros(ids.shape[0], np.int32) if with_mrca else None
        lib().oracle_distances_f64_climb(
            self.parent, self.node_depths(), self.distance, ids, ids.shape[0], out,
            m.ctypes.data if with_mrca else None,
        )
        return (out, m) if with_mrca else out


def pearson_f32(x, y):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    if len(x) != len(y):
        raise Exception("vectors must be the same length.", (len(x), len(y)))
    return float(lib().oracle_pearson_f32(x, y, len(x)))


def pearson_f64(x, y):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    return float(lib().oracle_pearson_f64(x, y, len(x)))


def linked_pairs(linklist):
    ll = np.ascontiguousarray(linklist, np.int64)
    L = ll.shape[0]
    size = L * (L - 1) // 2
    ids_a = np.empty((size, 2), np.int64)
    ids_b = np.empty((size, 2), np.int64)
    lib().oracle_linked_pairs(ll, L, ids_a, ids_b)
    return ids_a, ids_b


def random_ints(seed, n_links, count):
    """xorshift64* stream of MuchTree.pyx:2936-2949; returns (values, new_seed)."""
    s = C.c_uint64(int(seed))
    out = np.empty(count, np.uint64)
    f = lib().oracle_random_int
    for i in range(count):
        out[i] = f(C.byref(s), n_links)
    return out, int(s.value)


def sample_bucket(seed, linklist, n):
    ll = np.ascontiguousarray(linklist, np.int64)
    s = C.c_uint64(int(seed))
    qa = np.empty((n, 2), np.int64)
    qb = np.empty((n, 2), np.int64)
    lib().oracle_sample_bucket(C.byref(s), ll, ll.shape[0], n, qa, qb)
    return qa, qb, int(s.value)


# ---- per-clade scan: the reference's own loop, restated -------------------------
def get_leaves(left, right, node):
    """SuchTree.get_leaves (MuchTree.pyx:427-463): breadth-first queue from `node`."""
    to_visit = [int(node)]
    out = []
    for cur in to_visit:
        if left[cur] == -1:
            out.append(cur)
        else:
            to_visit.append(int(left[cur]))
            to_visit.append(int(right[cur]))
    return np.array(out, dtype=np.int64)


def subset_linklist(linklist, leafs, column):
    """What subset_b(node) / subset_a(node) + _build_linklist (MuchTree.pyx:2845-2898)
    leave of a link list: the rows whose id in `column` (0 = TreeB, 1 = TreeA) is one of
    `leafs`, relative order kept (the link list is built column by column either way)."""
    ll = np.ascontiguousarray(linklist, np.int64)
    return np.ascontiguousarray(ll[np.isin(ll[:, column], np.asarray(leafs, np.int64))])


def clade_scan(tree_a, tree_b, left, right, linklist, nodes, side):
    """for node in nodes: subset_b(node) [side 'b'] or subset_a(node) [side 'a'];
    linked_distances(); pearson() -- docs/examples/SuchLinkedTree_examples.md:286-310 --
    with O2 distances (fp64 path sums) and the fp64 pearson.  left/right: child arrays of
    the scanned tree.  Returns (n_links, r) per node; r = nan when there is no link pair."""
    column = 0 if side == "b" else 1
    n_links = np.zeros(len(nodes), np.int64)
    r = np.full(len(nodes), np.nan)
    for k, node in enumerate(nodes):
        sub = subset_linklist(linklist, get_leaves(left, right, node), column)
        n_links[k] = sub.shape[0]
        if sub.shape[0] < 2:
            continue
        ids_a, ids_b = linked_pairs(sub)  # :2918-2925
        r[k] = pearson_f64(tree_a.distances_f64(ids_a), tree_b.distances_f64(ids_b))
    return n_links, r
