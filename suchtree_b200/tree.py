"""SuchTree: host-side mirror of the reference class for the patristic-distance path.

Same constructor input, method names, argument meaning, return types and error
behaviour as the reference's `SuchTree` (MuchTree.pyx:90-2518) for every entry
point on the hot path; the computation itself happens in
libsuchtree_b200.so (hand-written sm_100a kernels) -- there is no CPU fallback.
"""
import ctypes as C
import os
import warnings
from numbers import Integral
from urllib.parse import urlparse

import numpy as np

from . import _lib, newick
from .exceptions import InvalidNodeError, NodeNotFoundError, TreeStructureError
from .extras import TreeExtras


def _deprecation_warning(old_name, new_name, version="2.0"):
    # same text as the reference's helper (MuchTree.pyx:33-42)
    warnings.warn(
        f"{old_name} is deprecated and will be removed in SuchTree {version}. Use {new_name} instead.",
        DeprecationWarning,
        stacklevel=3,
    )


def _read_tree_input(tree_input):
    """Input dispatch of MuchTree.pyx:138-155: URL, NEWICK string, or file path."""
    if not isinstance(tree_input, str):
        raise TypeError("tree_input must be a str (NEWICK text, file path or URL)")
    if urlparse(tree_input).scheme in ("http", "https", "ftp"):
        from urllib.request import urlopen

        return urlopen(tree_input).read().decode()
    if (
        "(" in tree_input
        and ")" in tree_input
        and tree_input.count("(") == tree_input.count(")")
        and tree_input.endswith(";")
    ):
        return tree_input
    if tree_input.endswith(".gz"):
        import gzip

        with gzip.open(tree_input, "rt") as f:
            return f.read()
    with open(tree_input) as f:
        return f.read()


class SuchTree(TreeExtras):
    """Immutable, strictly bifurcating phylogenetic tree resident on one B200.

    SuchTree(tree_input) accepts what the reference accepts (MuchTree.pyx:126-155).
    Extras, not in the reference: `device=` picks the GPU (default: LOCAL_RANK or
    0), and SuchTree.from_arrays() builds from node arrays directly.
    """

    def __init__(self, tree_input, device=None, _flat=None, _block_shift=0, _micro_shift=0, _wide=False):
        ft = _flat if _flat is not None else newick.flatten(_read_tree_input(tree_input))
        self._ft = ft
        self._epsilon = float(np.finfo(np.float64).eps)  # MuchTree.pyx:136
        self._leaves = ft.leaves
        self._leaf_nodes = None
        self._RED = {}
        self._rd = None
        self._leaf_ids = None
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self._handle = C.c_void_p()
        L = _lib.lib()
        parent = np.ascontiguousarray(ft.parent, np.int32)
        left = np.ascontiguousarray(ft.left, np.int32)
        right = np.ascontiguousarray(ft.right, np.int32)
        dist = np.ascontiguousarray(ft.distance, np.float32)
        rc = L.st_tree_create_ex(
            int(device), int(ft.size), parent.ctypes.data, left.ctypes.data, right.ctypes.data,
            dist.ctypes.data, int(_block_shift), int(_micro_shift), 1 if _wide else 0, C.byref(self._handle),
        )
        _lib.check(rc)
        info = _lib.TreeInfo()
        _lib.check(L.st_tree_get_info(self._handle, C.byref(info)))
        self._info = info
        self._size = int(info.n_nodes)
        self._depth = int(info.depth)
        self._n_leaves = int(info.n_leaves)
        self._root = int(info.root)
        self.device = int(info.device)

    @classmethod
    def from_arrays(cls, parent, left, right, distance, leaf_names=None, device=None, **kw):
        """Build from the reference's Node fields as arrays (ids must be in-order
        ranks; validated by the library).  leaf_names: optional names of the leaves
        in ascending id order."""
        ft = newick.FlatTree()
        ft.parent = np.asarray(parent, np.int32)
        ft.left = np.asarray(left, np.int32)
        ft.right = np.asarray(right, np.int32)
        ft.distance = np.asarray(distance, np.float32)
        ft.size = int(ft.parent.shape[0])
        ft.support = np.full(ft.size, -1.0, np.float32)
        leaf_ids = np.nonzero(ft.left == -1)[0]
        ft.n_leaves = int(leaf_ids.shape[0])
        ft.internal_nodes = np.nonzero(ft.left != -1)[0].astype(np.int64)
        r = np.nonzero(ft.parent == -1)[0]
        ft.root = int(r[0]) if r.size else -1
        ft.leaves = None if leaf_names is None else {str(nm): int(i) for nm, i in zip(leaf_names, leaf_ids)}
        return cls(None, device=device, _flat=ft, **kw)

    @classmethod
    def from_flat(cls, ft, device=None, **kw):
        return cls(None, device=device, _flat=ft, **kw)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().st_tree_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ====== properties (MuchTree.pyx:236-301) ======
    @property
    def size(self):
        return self._size

    @property
    def depth(self):
        return self._depth

    @property
    def num_leaves(self):
        return self._n_leaves

    @property
    def leaves(self):
        if self._leaves is None:  # synthetic trees: names are generated on first use
            ids = np.nonzero(self._ft.left == -1)[0]
            self._leaves = {"L%d" % (int(i) // 2): int(i) for i in ids}
            self._ft.leaves = self._leaves
        return self._leaves

    @property
    def leaf_nodes(self):
        if self._leaf_nodes is None:
            self._leaf_nodes = {v: k for k, v in self.leaves.items()}
        return self._leaf_nodes

    @property
    def root_node(self):
        return self._root

    @property
    def internal_nodes(self):
        return self._ft.internal_nodes

    @property
    def all_nodes(self):
        return np.concatenate((np.array(list(self.leaves.values())), np.array(list(self.internal_nodes))))

    @property
    def leaf_node_ids(self):
        if self._leaf_ids is None:  # the tree is immutable: built once, handed out as copies
            self._leaf_ids = np.array(list(self.leaves.values()))
        return self._leaf_ids.copy()

    @property
    def leaf_names(self):
        return list(self.leaves.keys())

    @property
    def polytomy_epsilon(self):
        return self._epsilon

    @polytomy_epsilon.setter
    def polytomy_epsilon(self, new_epsilon):
        self._epsilon = new_epsilon

    # deprecated aliases (MuchTree.pyx:2374-2414)
    @property
    def length(self):
        _deprecation_warning("length property", "size")
        return self.size

    @property
    def leafs(self):
        _deprecation_warning("leafs property", "leaves")
        return self.leaves

    @property
    def leafnodes(self):
        _deprecation_warning("leafnodes property", "leaf_nodes")
        return self.leaf_nodes

    @property
    def n_leafs(self):
        _deprecation_warning("n_leafs property", "num_leaves")
        return self.num_leaves

    @property
    def root(self):
        _deprecation_warning("root property", "root_node")
        return self.root_node

    @property
    def polytomy_distance(self):
        _deprecation_warning("polytomy_distance property", "polytomy_epsilon")
        return self.polytomy_epsilon

    @polytomy_distance.setter
    def polytomy_distance(self, value):
        _deprecation_warning("polytomy_distance property", "polytomy_epsilon")
        self.polytomy_epsilon = value

    # ====== validation helpers (MuchTree.pyx:2255-2300) ======
    def _validate_node(self, node):
        if isinstance(node, str):
            if node not in self.leaves:
                raise NodeNotFoundError(node)
            return self.leaves[node]
        if not isinstance(node, Integral):
            raise TypeError("Node must be int or str, got {t}".format(t=str(type(node))))
        node_id = int(node)
        if node_id < 0 or node_id >= self.size:
            raise InvalidNodeError(node_id, self.size)
        return node_id

    def _validate_nodes(self, nodes):
        """int64 ids of a sequence of nodes: one vectorised range check when they are all
        integers (the per-node loop costs ~1 us per id), else node by node as the reference."""
        if isinstance(nodes, np.ndarray) and nodes.dtype.kind in "iu" and nodes.ndim == 1:
            ids = nodes.astype(np.int64)
        elif isinstance(nodes, (list, tuple, range)) and all(type(v) is int for v in nodes):
            ids = np.fromiter(nodes, dtype=np.int64, count=len(nodes))
        else:
            return np.array([self._validate_node(nd) for nd in nodes], dtype=np.int64)
        if ids.size:
            bad = (ids < 0) | (ids >= self.size)
            if bad.any():
                raise InvalidNodeError(int(ids[np.argmax(bad)]), self.size)  # the first offender, as the loop
        return ids

    def _validate_node_pair(self, a, b):
        return self._validate_node(a), self._validate_node(b)

    # ====== host-side structure queries used around the path ======
    def get_parent(self, node):
        return int(self._ft.parent[self._validate_node(node)])

    def get_children(self, node):
        i = self._validate_node(node)
        return int(self._ft.left[i]), int(self._ft.right[i])

    def is_leaf(self, node):
        return bool(self._ft.left[self._validate_node(node)] == -1)

    def is_internal(self, node):
        return not self.is_leaf(node)

    def is_root(self, node):
        return self._validate_node(node) == self._root

    def _clade_interval(self, node_id):
        """[first, last] id of the clade under node_id (ids are in-order ranks, so a
        clade is a contiguous id interval)."""
        lo = hi = node_id
        left, right = self._ft.left, self._ft.right
        while left[lo] != -1:
            lo = left[lo]
        while right[hi] != -1:
            hi = right[hi]
        return int(lo), int(hi)

    def _clade_intervals(self):
        """(lo, hi) arrays: [first, last] id of the clade under every node, by pointer
        doubling over the left-most / right-most child maps (log2(depth) numpy passes,
        so the 10^6-deep caterpillar costs 20 of them)."""
        if getattr(self, "_clade_lo_hi", None) is None:
            ids = np.arange(self._size, dtype=np.int64)
            out = []
            for child in (self._ft.left, self._ft.right):
                f = np.where(np.asarray(child) == -1, ids, np.asarray(child, dtype=np.int64))
                while True:
                    g = f[f]
                    if np.array_equal(g, f):
                        break
                    f = g
                out.append(f)
            self._clade_lo_hi = (out[0], out[1])
        return self._clade_lo_hi

    def is_ancestor(self, a, b):
        """1 if a is an ancestor of b, -1 if b is an ancestor of a, else 0
        (MuchTree.pyx is_ancestor semantics)."""
        a, b = self._validate_node_pair(a, b)
        if a == b:
            return 0
        lo, hi = self._clade_interval(a)
        if lo <= b <= hi:
            return 1
        lo, hi = self._clade_interval(b)
        if lo <= a <= hi:
            return -1
        return 0

    def get_leaves(self, node):
        """Leaf ids below `node` in the reference's order (breadth-first queue,
        MuchTree.pyx:449-463), computed level by level."""
        cur = np.array([self._validate_node(node)], dtype=np.int64)
        left, right = self._ft.left, self._ft.right
        out = []
        while cur.size:
            l = left[cur]
            leaf = l == -1
            out.append(cur[leaf])
            inner = cur[~leaf]
            nxt = np.empty(2 * inner.size, dtype=np.int64)
            nxt[0::2] = left[inner]
            nxt[1::2] = right[inner]
            cur = nxt
        return np.concatenate(out) if out else np.empty(0, np.int64)

    def get_descendants(self, node):
        """Generator over the ids of `node` and everything below it, in the reference's
        order: the start node first, then breadth-first queue order (MuchTree.pyx:396-425)."""
        left, right = self._ft.left, self._ft.right
        to_visit = [self._validate_node(node)]
        for cur in to_visit:  # the list grows while it is iterated, as in the reference
            l = int(left[cur])
            if l != -1:
                to_visit.append(l)
                to_visit.append(int(right[cur]))
            yield int(cur)

    def traverse_preorder(self, from_node=None):
        """Node ids in preorder (root, left, right); MuchTree.pyx:1502-1536."""
        start = self._root if from_node is None else self._validate_node(from_node)
        left, right = self._ft.left, self._ft.right
        stack = [start]
        while stack:
            cur = stack.pop()
            r, l = int(right[cur]), int(left[cur])
            if r != -1:
                stack.append(r)
            if l != -1:
                stack.append(l)
            yield cur

    def pre_order(self):
        _deprecation_warning("pre_order()", "traverse_preorder()")
        return self.traverse_preorder()

    # ====== thin callers of the device index ======
    def _root_distances(self):
        """fp64 root distance of every node, as built on the device (hi + lo of the
        double-double); fetched once."""
        if self._rd is None:
            _, hi, lo = self.export_index()
            self._rd = hi + lo
        return self._rd

    def distance_to_root(self, node):
        """Distance from a node to the root; MuchTree.pyx:813-850.  The reference walks
        to the root adding fp32 edges; here it is one lookup in the device-built root
        distances (fp64 sum of the same fp32-quantised edges)."""
        node_id = self._validate_node(node)
        if self._has_minus_one_edge():
            # the reference's walk stops at the first edge whose length is exactly -1 (the
            # root's sentinel, MuchTree.pyx:845-849) -- also when a real edge has that length
            d, i = np.float32(0.0), node_id
            dist, parent = self._ft.distance, self._ft.parent
            while i != self._root and dist[i] != -1:
                d = np.float32(d + dist[i])
                i = int(parent[i])
            return float(d)
        return float(self._root_distances()[node_id])

    def _has_minus_one_edge(self):
        if getattr(self, "_minus_one_edge", None) is None:
            m = np.asarray(self._ft.distance) == -1
            m[self._root] = False
            self._minus_one_edge = bool(m.any())
        return self._minus_one_edge

    def get_distance_to_root(self, node):
        _deprecation_warning("get_distance_to_root()", "distance_to_root()")
        return self.distance_to_root(node)

    @property
    def relative_evolutionary_divergence(self):
        """RED of every node (Parks et al. 2018), dict node id -> RED in the reference's
        preorder insertion order; MuchTree.pyx:303-330.  a = edge to the parent, b = mean
        distance from the node to its leaf descendants.  The reference calls distance()
        once per (node, leaf) pair -- O(N x leaves); here b comes from one bottom-up
        pass, S[v] = S[l] + n[l] e[l] + S[r] + n[r] e[r] (all terms of one sign for
        ordinary trees: no cancellation, epsilon edges survive), level by level."""
        if not self._RED:
            n = self.size
            parent, left, right = self._ft.parent, self._ft.left, self._ft.right
            edge = self._ft.distance.astype(np.float64)
            edge[self._root] = 0.0
            depth, _, _ = self.export_index()  # device-built node depths
            levels = np.unique(depth)
            S = np.zeros(n, dtype=np.float64)
            cnt = (left == -1).astype(np.float64)
            for d in levels[::-1]:
                lvl = np.nonzero((depth == d) & (left != -1))[0]
                if lvl.size:
                    l, r = left[lvl], right[lvl]
                    S[lvl] = (S[l] + cnt[l] * edge[l]) + (S[r] + cnt[r] * edge[r])
                    cnt[lvl] = cnt[l] + cnt[r]
            b = S / cnt
            red = np.zeros(n, dtype=np.float64)
            for d in levels:
                if d == 0:
                    continue
                lvl = np.nonzero(depth == d)[0]
                ab = edge[lvl] + b[lvl]
                if np.any(ab == 0):
                    bad = int(lvl[np.nonzero(ab == 0)[0][0]])
                    raise Exception("node {n} : a={a}, b={b}".format(n=bad, a=edge[bad], b=b[bad]))
                P = red[parent[lvl]]
                red[lvl] = P + (edge[lvl] / ab) * (1 - P)
            out = {self._root: 0}
            for node in list(self.traverse_preorder())[1:]:
                out[node] = float(red[node])
            self._RED = out
        return self._RED

    RED = relative_evolutionary_divergence

    def relationships(self):
        """DataFrame of all leaf pairs with distance, root distances, MRCA and the two
        legs of the path; MuchTree.pyx:2158-2179 (pair orientation is random there:
        sample([a,b],2)).  One bulk distance launch + one bulk MRCA launch."""
        from itertools import combinations
        from random import sample

        import pandas as pd

        pairs = [sample([a, b], 2) for a, b in combinations(self.leaves.keys(), 2)]
        ids = np.array([(self.leaves[a], self.leaves[b]) for a, b in pairs], dtype=np.int64).reshape(-1, 2)
        rd = self._root_distances()
        if len(pairs):
            distances = self.distances_bulk(ids).tolist()
            mrca = self.common_ancestors_bulk(ids).astype(np.int64)
        else:
            distances, mrca = [], np.zeros(0, np.int64)
        a_to_root = rd[ids[:, 0]]
        b_to_root = rd[ids[:, 1]]
        mrca_to_root = rd[mrca]
        return pd.DataFrame({
            "a": [p[0] for p in pairs],
            "b": [p[1] for p in pairs],
            "distance": distances,
            "a_to_root": a_to_root.tolist(),
            "b_to_root": b_to_root.tolist(),
            "mrca": mrca.tolist(),
            "mrca_to_root": mrca_to_root.tolist(),
            "a_to_mrca": (a_to_root - mrca_to_root).tolist(),
            "b_to_mrca": (b_to_root - mrca_to_root).tolist(),
        })

    # ====== the hot path ======
    def distance(self, a, b):
        """Patristic distance between two nodes (ids or leaf names); MuchTree.pyx:852-870."""
        node_a, node_b = self._validate_node_pair(a, b)
        return float(self.distances_bulk(np.array([[node_a, node_b]], dtype=np.int64))[0])

    def distances_bulk(self, pairs, out=None):
        """Distances for an (n,2) array of node-id pairs; MuchTree.pyx:872-909.

        Accepts what the reference accepts: an int64 (n,2) ndarray with any strides,
        or anything np.array(..., dtype=int64) understands.  Wrong shape ->
        ValueError; non-int64 ndarray -> ValueError (the reference's memoryview
        'Buffer dtype mismatch'); out-of-range id -> InvalidNodeError(id, size)
        with the same id the reference reports (max id if >= size, else min id).
        """
        pairs = self._coerce_pairs(pairs)
        n = pairs.shape[0]
        if out is None:
            # a fresh array, as the reference returns (MuchTree.pyx:907); large ones come from
            # the library's page-locked pool so that the D2H copies land in them directly
            result = _lib.result_empty((n,), np.float64)
            if pairs.flags.c_contiguous:
                _lib.maybe_register(pairs)  # large inputs seen repeatedly are page-locked in place
        else:  # extension: caller-provided (e.g. pinned) float64 result buffer
            result = out
            if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (n,)
                    and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous float64 array of shape (n,)")
        if n:
            s0, s1 = pairs.strides[0] // 8, pairs.strides[1] // 8
            rc = _lib.lib().st_distances(self._handle, pairs.ctypes.data, s0, s1, n, result.ctypes.data)
            _lib.check(rc, self.size)
        return result

    def _coerce_pairs(self, pairs):
        if not isinstance(pairs, np.ndarray):
            pairs = np.array(pairs, dtype=np.int64)
        if pairs.ndim != 2 or pairs.shape[1] != 2:
            shape = str(pairs.shape[:2]) if pairs.ndim >= 2 else str(pairs.shape)
            raise ValueError("Expected (n, 2) array, got shape {shape}".format(shape=shape))
        if pairs.dtype != np.int64:
            raise ValueError(
                "Buffer dtype mismatch, expected 'long' but got '%s'" % pairs.dtype.name)
        if pairs.shape[0] == 0:
            # the reference fails in pairs.max() on an empty array (MuchTree.pyx:897)
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        if any(s % 8 for s in pairs.strides):
            pairs = np.ascontiguousarray(pairs)
        return pairs

    def distances(self, pairs):
        _deprecation_warning("distances()", "distances_bulk()")
        return self.distances_bulk(pairs)

    def distances_by_name(self, pairs):
        """MuchTree.pyx:945-979."""
        if not isinstance(pairs, list):
            raise TypeError("pairs must be a list of tuples")
        leaves = self.leaves
        # The name -> id walk is what bounds this entry point (the distances take ~1 ms per 10^6
        # pairs).  Fast paths, in order: the same walk as one C loop (pyglue/st_pynames.c, ~3.5x
        # faster than any Python form), else one C-level pass of dict lookups.  Anything unexpected
        # -- a missing name, a non-string, a pair that is not a pair -- falls through to the
        # reference's loop below, which raises the reference's error.
        ids = self._names_to_ids(pairs, 2)
        if ids is not None:
            return self.distances_bulk(ids).tolist() if len(ids) else []
        try:
            from itertools import chain

            if all(map(lambda p: len(p) == 2, pairs)):
                ids = np.fromiter(map(leaves.__getitem__, chain.from_iterable(pairs)), dtype=np.int64,
                                  count=2 * len(pairs)).reshape(-1, 2)
                if len(ids):  # every key of `leaves` is a str, so every element looked up was one
                    return self.distances_bulk(ids).tolist()
        except (KeyError, TypeError, ValueError):
            pass
        node_pairs = []
        for i, (name_a, name_b) in enumerate(pairs):
            if not isinstance(name_a, str) or not isinstance(name_b, str):
                raise TypeError("Pair {i}: both elements must be strings".format(i=str(i)))
            if name_a not in leaves:
                raise NodeNotFoundError(name_a)
            if name_b not in leaves:
                raise NodeNotFoundError(name_b)
            node_pairs.append((leaves[name_a], leaves[name_b]))
        if not node_pairs:
            return []
        return self.distances_bulk(np.array(node_pairs, dtype=np.int64)).tolist()

    def _names_to_ids(self, rows, width):
        """(n, width) int64 ids of a list of name tuples through the C glue, or None when the glue
        is unavailable or some row is not `width` leaf names (the caller's slow path decides how
        that fails)."""
        g = _lib.py_glue()
        if g is None or type(rows) is not list or type(self.leaves) is not dict:
            return None
        ids = np.empty((len(rows), width), dtype=np.int64)
        if g.st_py_names_to_ids(rows, self.leaves, ids.ctypes.data, width) != -1:
            return None
        return ids

    def common_ancestor(self, a, b):
        """MRCA node id; MuchTree.pyx:1128-1149."""
        node_a, node_b = self._validate_node_pair(a, b)
        return int(self.common_ancestors_bulk(np.array([[node_a, node_b]], dtype=np.int64))[0])

    def mrca(self, a, b):
        _deprecation_warning("mrca()", "common_ancestor()")
        return self.common_ancestor(a, b)

    def common_ancestors_bulk(self, pairs):
        """Batched MRCA ids (int32) for an (n,2) int64 array -- the bulk form of
        common_ancestor(); not in the reference API, same validation as distances_bulk."""
        pairs = self._coerce_pairs(pairs)
        n = pairs.shape[0]
        out = np.empty(n, dtype=np.int32)
        s0, s1 = pairs.strides[0] // 8, pairs.strides[1] // 8
        rc = _lib.lib().st_mrca(self._handle, pairs.ctypes.data, s0, s1, n, out.ctypes.data)
        _lib.check(rc, self.size)
        return out

    # ====== quartet topologies (MuchTree.pyx:1203-1421) ======
    def quartet_topologies_bulk(self, quartets):
        """(n,4) node ids in arbitrary order -> (n,4) int64 rows ordered so that
        {row[0],row[1]} and {row[2],row[3]} are the sister pairs; MuchTree.pyx:1271-1329.
        Six range-minimum MRCA lookups per quartet on the GPU."""
        if not isinstance(quartets, np.ndarray):
            quartets = np.array(quartets, dtype=np.int64)
        if quartets.ndim != 2 or quartets.shape[1] != 4:
            raise ValueError("Expected (n, 4) array, got shape {shape}".format(shape=quartets.shape))
        if quartets.dtype != np.int64:
            raise ValueError("Buffer dtype mismatch, expected 'long' but got '%s'" % quartets.dtype.name)
        if quartets.shape[0] == 0:
            # the reference fails in quartets.max() on an empty array (MuchTree.pyx:1303)
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        if any(s % 8 for s in quartets.strides):
            quartets = np.ascontiguousarray(quartets)
        n = quartets.shape[0]
        out = _lib.result_empty((n, 4), np.int64, zero_small=True)  # every row is written by the kernel
        s0, s1 = quartets.strides[0] // 8, quartets.strides[1] // 8
        rc = _lib.lib().st_quartet_topologies(self._handle, quartets.ctypes.data, s0, s1, n, out.ctypes.data)
        _lib.check(rc, self.size)
        return out

    def quartet_topology(self, a, b, c, d):
        """Topology of one quartet as a frozenset of two frozensets of sisters
        (MuchTree.pyx:1203-1247): the pairs whose MRCA is unique among the six pair
        MRCAs are sisters; leaf names come back when any input was a name."""
        from itertools import combinations

        nodes = [a, b, c, d]
        node_ids = [self._validate_node(node) for node in nodes]
        has_strings = any(isinstance(node, str) for node in nodes)
        pairs = [frozenset((x, y)) for x, y in combinations(node_ids, 2)]
        M = [int(m) for m in self.common_ancestors_bulk(
            np.array(list(combinations(node_ids, 2)), dtype=np.int64))]
        sisters = [pairs[M.index(i)] for i in M if M.count(i) == 1]
        if len(sisters) == 1:
            sisters.append(sisters[0] ^ frozenset(node_ids))
        if has_strings:
            (w, x), (y, z) = sisters
            return frozenset((frozenset((self.leaf_nodes[w], self.leaf_nodes[x])),
                              frozenset((self.leaf_nodes[y], self.leaf_nodes[z]))))
        return frozenset(sisters)

    def quartet_topologies_by_name(self, quartets):
        """MuchTree.pyx:1378-1421 (the later of the reference's two definitions)."""
        quartet_array = self._names_to_ids(quartets, 4)
        if quartet_array is None or not len(quartet_array):
            quartet_ids = []
            for i, (a, b, c, d) in enumerate(quartets):
                if not all(isinstance(name, str) for name in (a, b, c, d)):
                    raise TypeError(f"Quartet {i}: all elements must be strings")
                try:
                    quartet_ids.append([self.leaves[a], self.leaves[b], self.leaves[c], self.leaves[d]])
                except KeyError as e:
                    raise NodeNotFoundError(str(e).strip("'"))
            quartet_array = np.array(quartet_ids, dtype=np.int64)
        topologies = self.quartet_topologies_bulk(quartet_array)
        names = self.leaf_nodes
        return [frozenset((frozenset((names[a], names[b])), frozenset((names[c], names[d]))))
                for a, b, c, d in topologies.tolist()]

    # deprecated names (MuchTree.pyx:2461-2475)
    def get_quartet_topology(self, a, b, c, d):
        _deprecation_warning("get_quartet_topology()", "quartet_topology()")
        return self.quartet_topology(a, b, c, d)

    def quartet_topologies(self, quartets):
        _deprecation_warning("quartet_topologies()", "quartet_topologies_bulk()")
        return self.quartet_topologies_bulk(quartets)

    def quartet_topologies_device(self, d_quartets_ptr, n, d_out_ptr, stream=None, idx_bits=64):
        """Device-resident variant: contiguous int64 (idx_bits=64) or int32 (idx_bits=32)
        (n,4) in and out."""
        fn = _lib.lib().st_quartet_topologies_device if idx_bits == 64 else _lib.lib().st_quartet_topologies_device32
        rc = fn(self._handle, d_quartets_ptr, n, d_out_ptr, stream)
        _lib.check(rc, self.size)

    def nearest_neighbors(self, node, k=1, from_nodes=None):
        """MuchTree.pyx:1032-1080."""
        if k <= 0:
            raise ValueError("k must be positive")
        q = self._validate_node(node)
        if from_nodes is None:
            # all leaves (minus the query if it is one), as arrays: names are looked up
            # for the k winners only
            from_ids = np.asarray(self.leaf_node_ids, dtype=np.int64)
            if self.is_leaf(q):
                from_ids = from_ids[from_ids != q]
            from_orig = None
        else:
            from_ids = np.array([self._validate_node(n) for n in from_nodes], dtype=np.int64)
            from_orig = list(from_nodes)
        if from_ids.shape[0] == 0:
            return []
        pairs = np.empty((from_ids.shape[0], 2), dtype=np.int64)
        pairs[:, 0] = q
        pairs[:, 1] = from_ids
        d = self.distances_bulk(pairs)
        order = np.argsort(d)
        if from_orig is None:
            leaf_nodes = self.leaf_nodes
            return [(leaf_nodes[int(from_ids[i])], d[i]) for i in order[:k]]
        return [(from_orig[i], d[i]) for i in order[:k]]

    def pairwise_distances(self, nodes=None):
        """Symmetric (n,n) fp64 matrix; MuchTree.pyx:1082-1124.  The reference builds
        n(n-1)/2 Python tuples; here one tiled kernel writes the matrix."""
        if nodes is None:
            ids_ptr, n = None, self.num_leaves
            if self._leaves is not None:
                # the reference's default order is the `leaves` dict order (ascending id
                # unless duplicate names collapsed entries)
                ids = self.leaf_node_ids.astype(np.int64)
                if ids.shape[0] != n or not np.array_equal(ids, np.arange(0, 2 * n, 2)):
                    ids_ptr, n = ids, ids.shape[0]
        else:
            ids_ptr = self._validate_nodes(nodes)
            n = ids_ptr.shape[0]
        out = _lib.result_empty((n, n), np.float64, zero_small=True)  # every element is written by the kernels
        if n:
            rc = _lib.lib().st_distance_matrix(
                self._handle, None if ids_ptr is None else ids_ptr.ctypes.data, n, 0, n,
                out.ctypes.data, 0, None)
            _lib.check(rc, self.size)
        return out

    def distance_matrix(self, nodes=None):
        """MuchTree.pyx:1919-1956."""
        if nodes is None:
            node_ids = self.leaf_node_ids
            node_names = [self.leaf_nodes[nid] for nid in node_ids]
        else:
            node_ids = np.array([self._validate_node(nd) for nd in nodes])
            node_names = []
            for nid in node_ids:
                if self._ft.left[nid] == -1:
                    node_names.append(self.leaf_nodes[nid])
                else:
                    node_names.append(f"node_{nid}")
        return {
            "distance_matrix": self.pairwise_distances(nodes),
            "node_ids": node_ids,
            "node_names": node_names,
        }

    # ====== device-resident API (throughput path; not in the reference) ======
    def distances_device(self, d_pairs_ptr, n, d_out_ptr, idx_bits=32, d_mrca_ptr=None, stream=None):
        """Launch the query kernel on device pointers (ints); asynchronous on `stream`
        (a cudaStream_t as int, None = legacy default stream)."""
        rc = _lib.lib().st_distances_device(
            self._handle, d_pairs_ptr, idx_bits, n, d_out_ptr, d_mrca_ptr, stream)
        _lib.check(rc, self.size)

    def check_range(self, stream=None):
        _lib.check(_lib.lib().st_check_range(self._handle, stream), self.size)

    def random_leaf_pairs_device(self, seed, first_pair, n, d_pairs_ptr, idx_bits=32, stream=None):
        rc = _lib.lib().st_random_leaf_pairs_device(
            self._handle, seed, first_pair, n, d_pairs_ptr, idx_bits, stream)
        _lib.check(rc, self.size)

    def export_index(self):
        """(depth int32[n], rd_hi, rd_lo float64[n]) as built on the device."""
        n = self.size
        depth = np.empty(n, np.int32)
        hi = np.empty(n, np.float64)
        lo = np.empty(n, np.float64)
        _lib.check(_lib.lib().st_tree_export(self._handle, depth.ctypes.data, hi.ctypes.data, lo.ctypes.data))
        return depth, hi, lo

    @property
    def index_info(self):
        i = self._info
        return {k: getattr(i, k) for k, _ in i._fields_}
