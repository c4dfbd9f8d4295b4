"""SuchLinkedTrees and pearson(): host-side mirror of the reference's co-phylogeny
entry points on the hot path:

    SuchLinkedTrees(tree_a, tree_b, link_matrix)     MuchTree.pyx:2562-2667
    .linklist / subset_a / subset_b                  :2839-2898
    .linked_distances()                              :2900-2934
    .sample_linked_distances(sigma, buckets, n, maxcycles)   :2951-3079
    pearson(x, y)                                    :81-87

Link bookkeeping stays on the host (small, pandas-driven); every distance, the
sampler's random stream and the moment reductions run in libsuchtree_b200.so.
"""
import ctypes as C
import os

import numpy as np

from . import _lib, shard
from .extras import LinkedExtras
from .tree import SuchTree

_UINT64_MAX = np.iinfo(np.uint64).max


def pearson(x, y, device=None):
    """Pearson correlation of two float64 vectors; MuchTree.pyx:81-87 (the formula of :62-79
    with fp64 accumulators on the GPU: the vectors are streamed to the device in chunks and
    folded into shifted moments in one pass).  device: default LOCAL_RANK (one process per
    GPU), like SuchTree."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    x = _as_f64_vector(x)
    y = _as_f64_vector(y)
    if not len(x) == len(y):
        raise Exception("vectors must be the same length.", (len(x), len(y)))
    r = C.c_double(0.0)
    rc = _lib.lib().st_pearson(int(device), x.ctypes.data, y.ctypes.data, len(x), C.byref(r))
    _lib.check(rc)
    return float(r.value)


def _as_f64_vector(v):
    a = np.asarray(v)
    if a.dtype != np.float64:
        # the reference takes `double[:]` memoryviews
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % a.dtype.name)
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    return np.ascontiguousarray(a)


class SuchLinkedTrees(LinkedExtras):
    def __init__(self, tree_a, tree_b, link_matrix):
        # the reference seeds its xorshift64* state in __cinit__ (MuchTree.pyx:2572-2573),
        # i.e. before anything else happens -- same draw, same place
        self._seed = int(np.random.randint(_UINT64_MAX >> 1))
        self._TreeA = self._as_tree(tree_a)
        self._TreeB = self._as_tree(tree_b)
        TA, TB = self._TreeA, self._TreeB
        if TA.device != TB.device:
            raise Exception("both trees must live on the same device", (TA.device, TB.device))
        if not link_matrix.shape == (TA.num_leaves, TB.num_leaves):
            raise Exception("link_matrix shape must match tree leaf counts")
        if not set(link_matrix.axes[0]) == set(TA.leaves.keys()):
            raise Exception("axis[0] does not match TreeA leaf names")
        if not set(link_matrix.axes[1]) == set(TB.leaves.keys()):
            raise Exception("axis[1] does not match TreeB leaf names")

        self._row_ids = np.array(list(TA.leaves.values()))
        self._col_ids = np.array(list(TB.leaves.values()))
        self._row_names = list(TA.leaves.keys())
        self._col_names = list(TB.leaves.keys())
        self._n_rows = TA.num_leaves
        self._n_cols = TB.num_leaves
        self._row_map = np.zeros(TA.size, dtype=np.int64)
        self._row_map[self._row_ids] = np.arange(len(self._row_ids))
        self._col_of_leaf = np.full(TB.size, -1, dtype=np.int64)  # stands in for link_leaf (:2639)
        self._col_of_leaf[self._col_ids] = np.arange(len(self._col_ids))

        # link table (:2633-2650): column i = TreeB leaf col_names[i]; its links are the
        # TreeA leaf ids of the rows with value > 0, in the DataFrame's row order
        values = link_matrix.reindex(columns=self._col_names).to_numpy()
        row_leaf_ids = np.array([TA.leaves[name] for name in link_matrix.index], dtype=np.int64)
        cols, rows = np.nonzero((values > 0).T)  # column-major: by column, then DataFrame row order
        self._link_cols = cols.astype(np.int64)
        self._link_a = row_leaf_ids[rows]
        self._n_links = int(cols.shape[0])

        self._set_default_subset()

    @classmethod
    def from_linklist(cls, tree_a, tree_b, linklist):
        """Extra, not in the reference: the same object from link rows [TreeB leaf id,
        TreeA leaf id] (the layout of .linklist, :2869-2870) instead of a DataFrame --
        synthetic workloads with 10^5+ links, where a dense link matrix would not fit.
        Links are kept in the reference's order: by column (TreeB leaf, ascending id),
        then in the order given."""
        self = cls.__new__(cls)
        self._seed = int(np.random.randint(_UINT64_MAX >> 1))
        self._TreeA, self._TreeB = TA, TB = cls._as_tree(tree_a), cls._as_tree(tree_b)
        if TA.device != TB.device:
            raise Exception("both trees must live on the same device", (TA.device, TB.device))
        ll = np.ascontiguousarray(linklist, dtype=np.int64)
        if ll.ndim != 2 or ll.shape[1] != 2:
            raise ValueError("linklist must have shape (n_links, 2)")
        self._row_ids = TA.leaf_node_ids
        self._col_ids = TB.leaf_node_ids
        self._row_names = TA.leaf_names
        self._col_names = TB.leaf_names
        self._n_rows, self._n_cols = TA.num_leaves, TB.num_leaves
        self._row_map = np.full(TA.size, -1, dtype=np.int64)
        self._row_map[self._row_ids] = np.arange(self._n_rows)
        self._col_of_leaf = np.full(TB.size, -1, dtype=np.int64)
        self._col_of_leaf[self._col_ids] = np.arange(self._n_cols)
        for col, T, what in ((0, TB, self._col_of_leaf), (1, TA, self._row_map)):
            ids = ll[:, col]
            bad = (ids < 0) | (ids >= T.size)
            if not bad.any():
                bad = what[ids] < 0
            if bad.any():
                raise Exception("linklist: not a leaf id of Tree%s" % "BA"[col], int(ids[np.argmax(bad)]))
        cols = self._col_of_leaf[ll[:, 0]]
        perm = np.argsort(cols, kind="stable")
        self._link_cols = cols[perm]
        self._link_a = ll[perm, 1]
        self._n_links = int(ll.shape[0])
        self._set_default_subset()
        return self

    def _set_default_subset(self):
        # default subset = everything (:2652-2662)
        TA, TB = self._TreeA, self._TreeB
        self._subset_a_root = TA.root_node
        self._subset_b_root = TB.root_node
        self._subset_a_leafs = self._row_ids
        self._subset_b_leafs = self._col_ids
        self._subset_columns = np.arange(self._n_cols)
        self._subset_rows = np.arange(self._n_rows)
        self._subset_a_size = self._n_rows
        self._subset_b_size = self._n_cols
        self._np_linklist = np.ndarray((self._n_links, 2), dtype=np.int64)
        self._build_linklist()

    @staticmethod
    def _as_tree(t):
        if isinstance(t, str):
            return SuchTree(t)
        if type(t) == SuchTree:
            return t
        raise Exception("unknown input for tree", type(t))

    # ---- link list (:2845-2874) ------------------------------------------
    def _build_linklist(self):
        """rows [TreeB leaf id, TreeA leaf id], ordered by subset column, then by the
        link order inside the column, restricted to subset_a's leaves."""
        self._subset_version = getattr(self, "_subset_version", 0) + 1
        in_a = np.zeros(self._TreeA.size, dtype=bool)
        in_a[np.asarray(self._subset_a_leafs, dtype=np.int64)] = True
        # stable selection of the links of each subset column, in subset-column order
        order_of_col = np.full(self._n_cols, -1, dtype=np.int64)
        sc = np.asarray(self._subset_columns, dtype=np.int64)
        order_of_col[sc] = np.arange(sc.shape[0])
        keep = (order_of_col[self._link_cols] >= 0) & in_a[self._link_a]
        cols, a = self._link_cols[keep], self._link_a[keep]
        perm = np.argsort(order_of_col[cols], kind="stable")
        cols, a = cols[perm], a[perm]
        k = cols.shape[0]
        self._np_linklist[:k, 0] = self._col_ids[cols]
        self._np_linklist[:k, 1] = a
        self._subset_n_links = int(k)

    def subset_b(self, node_id):
        """:2876-2886"""
        if node_id > self._TreeB.size or node_id < 0:
            raise Exception("Node ID out of bounds.", node_id)
        self._subset_b_leafs = self._TreeB.get_leaves(node_id)
        self._subset_columns = self._col_of_leaf[self._subset_b_leafs]
        self._subset_b_size = len(self._subset_columns)
        self._subset_b_root = node_id
        self._build_linklist()

    def subset_a(self, node_id):
        """:2888-2898"""
        if node_id > self._TreeA.size or node_id < 0:
            raise Exception("Node ID out of bounds.", node_id)
        self._subset_a_leafs = self._TreeA.get_leaves(node_id)
        self._subset_rows = self._row_map[self._subset_a_leafs]
        self._subset_a_size = len(self._subset_rows)
        self._subset_a_root = node_id
        self._build_linklist()

    # ---- properties (:2681-2772, :2839-2843) --------------------------------
    TreeA = property(lambda self: self._TreeA)
    TreeB = property(lambda self: self._TreeB)
    n_links = property(lambda self: self._n_links)
    n_cols = property(lambda self: self._n_cols)
    n_rows = property(lambda self: self._n_rows)
    col_ids = property(lambda self: self._col_ids)
    row_ids = property(lambda self: self._row_ids)
    col_names = property(lambda self: self._col_names)
    row_names = property(lambda self: self._row_names)
    subset_columns = property(lambda self: self._subset_columns)
    subset_rows = property(lambda self: self._subset_rows)
    subset_a_leafs = property(lambda self: self._subset_a_leafs)
    subset_b_leafs = property(lambda self: self._subset_b_leafs)
    subset_a_size = property(lambda self: self._subset_a_size)
    subset_b_size = property(lambda self: self._subset_b_size)
    subset_a_root = property(lambda self: self._subset_a_root)
    subset_b_root = property(lambda self: self._subset_b_root)
    subset_n_links = property(lambda self: self._subset_n_links)

    # ---- per-column access and the dense matrix (:2774-2837) ------------------
    def _col_index(self, col):
        col_id = self._col_names.index(col) if isinstance(col, str) else col
        if col_id > self._n_cols:
            raise Exception("col_id out of bounds", col_id)
        return col_id

    def get_column_leafs(self, col, as_row_ids=False):
        """TreeA leaf ids (or row indices) linked to one column (a TreeB leaf), in the
        link matrix's row order; :2774-2792."""
        col_id = self._col_index(col)
        links = self._link_a[self._link_cols == col_id]
        return self._row_map[links] if as_row_ids else links.copy()

    def get_column_links(self, col):
        """Boolean row mask of one column; :2794-2809."""
        column = np.zeros(self._n_rows, dtype=bool)
        column[self._row_map[self._link_a[self._link_cols == self._col_index(col)]]] = True
        return column

    @property
    def linkmatrix(self):
        """Dense boolean (subset_a_size, subset_b_size) link matrix, built on access
        (:2811-2837; indexed by row id / column id exactly like the reference, including
        its behaviour under subsetting, which upstream marks FIXME)."""
        table = np.zeros((self._subset_a_size, self._subset_b_size), dtype=bool)
        in_a = np.zeros(self._TreeA.size, dtype=bool)
        in_a[np.asarray(self._subset_a_leafs, dtype=np.int64)] = True
        in_cols = np.zeros(self._n_cols, dtype=bool)
        in_cols[np.asarray(self._subset_columns, dtype=np.int64)] = True
        keep = in_cols[self._link_cols] & in_a[self._link_a]
        table[self._row_map[self._link_a[keep]], self._link_cols[keep]] = True
        return table

    @property
    def linklist(self):
        return self._np_linklist[: self._subset_n_links, :]

    def _linklist_c(self):
        return np.ascontiguousarray(self.linklist, dtype=np.int64)

    def _links(self):
        """Device-resident link list (st_links handle) of the current subset; rebuilt only
        when subset_a / subset_b changed the list."""
        ver = getattr(self, "_subset_version", 0)
        cur = getattr(self, "_links_handle", None)
        if cur is None or cur[0] != ver:
            self._drop_links()
            h = C.c_void_p()
            ll = self._linklist_c()
            _lib.check(_lib.lib().st_links_create(self._TreeA._handle, self._TreeB._handle, ll.ctypes.data,
                                                  self._subset_n_links, C.byref(h)))
            self._links_handle = (ver, h)
        return self._links_handle[1]

    def _drop_links(self):
        cur = getattr(self, "_links_handle", None)
        if cur is not None and cur[1].value:
            try:
                _lib.lib().st_links_destroy(cur[1])
            except Exception:
                pass
        self._links_handle = None

    def __del__(self):
        self._drop_links()
        self._drop_scan_handle()

    # ---- linked_distances (:2900-2934) ---------------------------------------
    def linked_distances(self):
        L = self._subset_n_links
        size = (L * (L - 1)) // 2
        ids_a = np.ndarray((size, 2), dtype=np.int64)
        ids_b = np.ndarray((size, 2), dtype=np.int64)
        out_a = np.zeros(size, dtype=np.float64)
        out_b = np.zeros(size, dtype=np.float64)
        if size:
            rc = _lib.lib().st_links_linked_distances(
                self._links(), out_a.ctypes.data, out_b.ctypes.data, ids_a.ctypes.data, ids_b.ctypes.data)
            _lib.check(rc)
        return {
            "TreeA": out_a,
            "TreeB": out_b,
            "ids_A": ids_a,
            "ids_B": ids_b,
            "n_pairs": size,
            "n_samples": size,
            "deviation_a": None,
            "deviation_b": None,
        }

    # ---- sample_linked_distances (:2951-3079) --------------------------------
    def sample_linked_distances(self, sigma=0.001, buckets=64, n=4096, maxcycles=100):
        """Bucketed sampling with the reference's convergence rule.  Link pairs are
        drawn with the reference's own xorshift64* stream (reproduced exactly on the
        device by jump-ahead), so for the same numpy seed the returned distances are
        the reference's, sample for sample."""
        sigma = float(np.float32(sigma))  # `float sigma` argument
        buckets, n, maxcycles = int(buckets), int(n), int(maxcycles)
        L = self._subset_n_links
        links = self._links()
        sums_a = np.zeros(buckets)
        sums_b = np.zeros(buckets)
        sumsq_a = np.zeros(buckets)
        sumsq_b = np.zeros(buckets)
        samples = 0
        # every cycle's distances land directly in the result arrays (the library stages the D2H
        # copies through its lane's page-locked buffer).  As the reference does (:2990-2995), room
        # for maxcycles cycles is reserved up front -- np.empty touches no pages, so only the cycles
        # that run cost memory -- up to 1 GiB per array; beyond that the arrays grow geometrically.
        # The result is a view of the first `cycles` cycles.
        bn = buckets * n
        cap = max(1, min(maxcycles, (1 << 27) // max(bn, 1)))
        out_a = np.empty(cap * bn)
        out_b = np.empty(cap * bn)
        seed = C.c_uint64(self._seed)
        cycles = 0
        lib = _lib.lib()
        while True:
            if cycles == cap:
                cap = min(max(maxcycles, cap + 1), 2 * cap)
                grown_a = np.empty(cap * bn)
                grown_b = np.empty(cap * bn)
                grown_a[: cycles * bn] = out_a[: cycles * bn]
                grown_b[: cycles * bn] = out_b[: cycles * bn]
                out_a, out_b = grown_a, grown_b
            rc = lib.st_links_sample_cycle(
                links, C.byref(seed), buckets, n,
                out_a.ctypes.data + cycles * bn * 8, out_b.ctypes.data + cycles * bn * 8,
                sums_a.ctypes.data, sumsq_a.ctypes.data, sums_b.ctypes.data, sumsq_b.ctypes.data)
            self._seed = int(seed.value)
            _lib.check(rc)
            samples += n
            with np.errstate(invalid="ignore"):
                dev_a = (sumsq_a / samples - (sums_a / samples) ** 2) ** 0.5
                dev_b = (sumsq_b / samples - (sums_b / samples) ** 2) ** 0.5
            # std of the bucket stds, with the reference's fp32 scalars (:3016-3019, :3057-3067)
            f32 = np.float32
            deviation_a = f32(0)
            deviation_b = f32(0)
            ssq_a = f32(0)
            ssq_b = f32(0)
            for i in range(buckets):
                deviation_a = f32(float(deviation_a) + dev_a[i])
                deviation_b = f32(float(deviation_b) + dev_b[i])
                ssq_a = f32(float(ssq_a) + dev_a[i] ** 2)
                ssq_b = f32(float(ssq_b) + dev_b[i] ** 2)
            with np.errstate(invalid="ignore"):
                deviation_a = f32((float(ssq_a / f32(buckets)) - float(deviation_a / f32(buckets)) ** 2) ** 0.5)
                deviation_b = f32((float(ssq_b / f32(buckets)) - float(deviation_b / f32(buckets)) ** 2) ** 0.5)
            cycles += 1
            if deviation_a < sigma and deviation_b < sigma:
                break
            if cycles >= maxcycles:
                return None
        return {
            "TreeA": out_a[: cycles * bn],
            "TreeB": out_b[: cycles * bn],
            "n_pairs": (L * (L - 1)) / 2,
            "n_samples": n * buckets * cycles,
            "deviation_a": float(deviation_a),
            "deviation_b": float(deviation_b),
        }

    # ---- throughput path (not in the reference API) --------------------------
    def sample_moments(self, n_samples, seed=0, first_sample=0, x0=0.0, y0=0.0, comm=None):
        """Philox-sampled link pairs, moments fused on the device; returns the
        _lib.Moments struct.  comm: a shard.MomentComm (one process per GPU) -- the six sums
        are then all-reduced by NCCL on the kernel's stream inside the call, and the struct
        holds the moments of ALL ranks' samples (every rank must call, same x0 / y0)."""
        m = _lib.Moments()
        rc = _lib.lib().st_links_sample_moments(
            self._links(), int(seed), int(first_sample), int(n_samples), float(x0), float(y0),
            None if comm is None else comm.handle, C.byref(m))
        _lib.check(rc)
        return m

    def linked_moments(self, first_pair=0, n_pairs=None, x0=0.0, y0=0.0, comm=None):
        """Moments of (d_A, d_B) over the link pairs [first_pair, first_pair + n_pairs) of
        linked_distances()'s enumeration, fused on the device (nothing materialised).
        comm: as in sample_moments()."""
        L = self._subset_n_links
        total = (L * (L - 1)) // 2
        if n_pairs is None:
            n_pairs = total - first_pair
        m = _lib.Moments()
        rc = _lib.lib().st_links_linked_moments(
            self._links(), int(first_pair), int(n_pairs), float(x0), float(y0),
            None if comm is None else comm.handle, C.byref(m))
        _lib.check(rc)
        return m

    def linked_pearson(self):
        """pearson(linked_distances()['TreeA'], linked_distances()['TreeB']) in one fused
        pass -- the inner step of the reference's per-clade correlation scans
        (subset_b(clade); linked_distances(); pearson())."""
        L = self._subset_n_links
        if L < 2:
            return 0.0
        # shift by the first pair's distances for conditioning
        ll = self.linklist
        x0 = self._TreeA.distance(int(ll[0, 1]), int(ll[1, 1]))
        y0 = self._TreeB.distance(int(ll[0, 0]), int(ll[1, 0]))
        return moments_pearson(self.linked_moments(x0=x0, y0=y0))

    def clade_moments(self, nodes=None, side="b", min_links=2, max_links=None):
        """The reference's per-clade scan -- for node in nodes: subset_b(node) (or
        subset_a for side="a"); linked_distances(); pearson()
        (docs/examples/SuchLinkedTree_examples.md:286-310) -- in one launch sequence:
        a clade is a contiguous id interval, so every subset is a run of the link list
        sorted by the scanned side's leaf id.  The subset currently set on the OTHER side
        stays in force (as it does across the reference's subset_b calls); the scanned
        side's own subset is replaced per clade, exactly as subset_b(node) replaces it.
        nodes: node ids of the scanned tree (default: all its internal nodes).  Clades
        with fewer than min_links or more than max_links links are counted but not
        computed (the example's `if SLT.subset_n_links < 10: continue`).
        Returns (node_ids, n_leafs, n_links, moments) with moments a float64 array
        (n_clades, 8) whose rows are laid out as the C struct st_moments -- n, x0, y0, sx, sy,
        sxx, syy, sxy; n = 0 where skipped -- see as_moments()."""
        if side not in ("a", "b"):
            raise ValueError("side must be 'a' or 'b'")
        T = self._TreeB if side == "b" else self._TreeA
        if nodes is None:
            nodes = np.nonzero(np.asarray(T._ft.left) != -1)[0]
        nodes = np.ascontiguousarray(nodes, dtype=np.int64).reshape(-1)
        if nodes.size and (nodes.min() < 0 or nodes.max() >= T.size):
            raise Exception("Node ID out of bounds.", int(nodes.max() if nodes.max() >= T.size else nodes.min()))
        lo_all, hi_all = T._clade_intervals()
        n = int(nodes.shape[0])
        big = n * 64 >= _lib.PINNED_RESULT_MIN_BYTES  # page-locked arrays: the copies are plain DMA
        new = _lib.pinned_empty if big else np.empty
        lo, hi = new((n,), np.int64), new((n,), np.int64)
        np.take(lo_all, nodes, out=lo)
        np.take(hi_all, nodes, out=hi)
        links = self._scan_handle(side)
        moments = new((n, 8), np.float64)  # rows laid out as st_moments; every row is written
        n_links = new((n,), np.int64)
        if n:
            rc = _lib.lib().st_links_clade_moments(
                links, 0 if side == "b" else 1, lo.ctypes.data, hi.ctypes.data, n, int(min_links),
                -1 if max_links is None else int(max_links), moments.ctypes.data, n_links.ctypes.data)
            _lib.check(rc)
        return nodes, (hi - lo) // 2 + 1, n_links, moments

    def _links_for_scan(self, side):
        """Link rows [TreeB id, TreeA id] under the OTHER side's current subset, in link-table
        order (by column, then the link matrix's row order: _build_linklist's order,
        :2845-2874); kept until a subset changes."""
        key = (side, self._subset_version)
        if getattr(self, "_scan_links", None) is None or self._scan_links[0] != key:
            if side == "b":
                in_a = np.zeros(self._TreeA.size, dtype=bool)
                in_a[np.asarray(self._subset_a_leafs, dtype=np.int64)] = True
                keep = in_a[self._link_a]
            else:
                in_cols = np.zeros(self._n_cols, dtype=bool)
                in_cols[np.asarray(self._subset_columns, dtype=np.int64)] = True
                keep = in_cols[self._link_cols]
            ll = np.empty((int(keep.sum()), 2), dtype=np.int64)
            ll[:, 0] = self._col_ids[self._link_cols[keep]]
            ll[:, 1] = self._link_a[keep]
            self._scan_links = (key, ll)
        return self._scan_links[1]

    def _scan_handle(self, side):
        """st_links handle of _links_for_scan(side): the sorted links and prefix counts the scan
        needs stay on the device until a subset changes."""
        ll = self._links_for_scan(side)
        key = self._scan_links[0]
        cur = getattr(self, "_scan_links_handle", None)
        if cur is None or cur[0] != key:
            self._drop_scan_handle()
            h = C.c_void_p()
            _lib.check(_lib.lib().st_links_create(self._TreeA._handle, self._TreeB._handle, ll.ctypes.data,
                                                  int(ll.shape[0]), C.byref(h)))
            self._scan_links_handle = (key, h)
        return self._scan_links_handle[1]

    def _drop_scan_handle(self):
        cur = getattr(self, "_scan_links_handle", None)
        if cur is not None and cur[1].value:
            try:
                _lib.lib().st_links_destroy(cur[1])
            except Exception:
                pass
        self._scan_links_handle = None

    def clade_pearson(self, nodes=None, side="b", min_links=2, max_links=None, rank=None, world=None):
        """Pearson r of (TreeA distance, TreeB distance) over all link pairs of every
        clade: dict of arrays node_ids, n_leafs (subset_b_size), n_links (subset_n_links),
        n_pairs, r (nan where the clade was skipped by min_links / max_links).
        rank / world (one process per GPU): this rank computes -- and returns the rows of --
        its share of the clades only, dealt out by link-pair count (shard.balanced_shares);
        the shares of all ranks partition `nodes`, no collective is involved."""
        if world is not None and int(world) > 1:
            # counts only (nothing is eligible with max_links < min_links), then this rank's share
            nodes, _, n_links, _ = self.clade_moments(nodes, side, 2, 1)
            ok = n_links >= max(int(min_links), 2)
            if max_links is not None:
                ok &= n_links <= int(max_links)
            weights = np.where(ok, n_links * (n_links - 1.0) * 0.5, 0.0)
            nodes = nodes[shard.balanced_shares(weights, int(world))[int(rank)]]
        nodes, n_leafs, n_links, m = self.clade_moments(nodes, side, min_links, max_links)
        cnt = np.ascontiguousarray(m[:, 0])
        done = np.nonzero(cnt > 0)[0]
        r = np.full(nodes.shape[0], np.nan)
        md = m[done]  # n, x0, y0, sx, sy, sxx, syy, sxy of the clades that were computed
        c = md[:, 0]
        cxx = md[:, 5] - md[:, 3] * md[:, 3] / c
        cyy = md[:, 6] - md[:, 4] * md[:, 4] / c
        cxy = md[:, 7] - md[:, 3] * md[:, 4] / c
        with np.errstate(invalid="ignore"):
            r[done] = cxy / np.sqrt(cxx * cyy + 1.0e-20)  # st_moments_pearson, MuchTree.pyx:79
        return {"node_ids": nodes, "n_leafs": n_leafs, "n_links": n_links,
                "n_pairs": cnt.astype(np.int64), "r": r}

    def sample_pearson(self, n_samples, seed=0, comm=None):
        """Sampled two-tree Pearson r over n_samples link pairs drawn with replacement.
        With comm (shard.MomentComm, one process per GPU) n_samples is the WHOLE job's
        sample count: every rank draws its own contiguous share of the Philox stream, the
        moments are all-reduced inside the call and every rank returns the same r."""
        if comm is None:
            return moments_pearson(self.sample_moments(n_samples, seed=seed))
        b, e = shard.pair_range(comm.rank, comm.world, int(n_samples))
        return moments_pearson(self.sample_moments(e - b, seed=seed, first_sample=b, comm=comm))


def as_moments(row):
    """One row of clade_moments()'s array as the _lib.Moments struct."""
    return _lib.Moments(*(float(v) for v in row))


def moments_pearson(m):
    return float(_lib.lib().st_moments_pearson(C.byref(m)))
