"""Seeded synthetic trees, generated straight into the reference's node layout
(in-order ids, leaves even / internal odd; MuchTree.pyx:171-216) -- the
workloads of BASELINE.json: random binary (Yule), balanced and caterpillar trees.

Edge lengths are fp32 uniform in [0.5, 1): with that exponent range fp64 prefix
sums of even a 10^6-deep path are exact (24 + 20 + 1 bits), so the fp64 oracle
and the kernels agree bit for bit.

Everything is vectorised level by level (no Python loop over nodes) except the
caterpillar, which is closed-form.
"""
import numpy as np

from .newick import FlatTree


def _edges(n_nodes, root, seed):
    rng = np.random.default_rng(seed)
    e = (0.5 + 0.5 * rng.random(n_nodes, dtype=np.float32)).astype(np.float32)
    e[e >= 1.0] = np.float32(0.75)
    e[root] = -1.0
    return e


def _finish(parent, left, right, root, seed, names):
    ft = FlatTree()
    n = parent.shape[0]
    ft.size = n
    ft.parent, ft.left, ft.right = parent, left, right
    ft.root = int(root)
    ft.distance = _edges(n, root, seed)
    ft.support = np.full(n, -1.0, np.float32)
    ft.n_leaves = (n + 1) // 2
    ft.internal_nodes = np.arange(1, n, 2, dtype=np.int64)
    if names:
        ft.leaves = {"L%d" % k: 2 * k for k in range(ft.n_leaves)}
    else:
        ft.leaves = None  # built lazily by SuchTree when asked for
    return ft


def _split_tree(n_leaves, split_fn, seed, names):
    """Top-down, level-synchronous construction.  A clade is a half-open leaf-rank
    interval [lo, hi); splitting it after k leaves puts its root at in-order id
    2*(lo+k)-1, so ids never need renumbering."""
    n = 2 * n_leaves - 1
    parent = np.full(n, -1, np.int32)
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    lo = np.array([0], np.int64)
    hi = np.array([n_leaves], np.int64)
    par = np.array([-1], np.int64)
    is_left = np.array([False])
    root = None
    while lo.size:
        size = hi - lo
        leaf = size == 1
        # leaves of this level
        lid = 2 * lo[leaf]
        parent[lid] = par[leaf]
        pl = par[leaf]
        il = is_left[leaf]
        ok = pl >= 0
        left[pl[ok & il]] = lid[ok & il]
        right[pl[ok & ~il]] = lid[ok & ~il]
        if root is None and leaf.any() and par[leaf][0] == -1:
            root = int(lid[0])
        # internal clades
        inn = ~leaf
        if not inn.any():
            break
        lo_i, hi_i, par_i, il_i = lo[inn], hi[inn], par[inn], is_left[inn]
        k = split_fn(hi_i - lo_i)
        mid = lo_i + k
        nid = 2 * mid - 1
        parent[nid] = par_i
        ok = par_i >= 0
        left[par_i[ok & il_i]] = nid[ok & il_i]
        right[par_i[ok & ~il_i]] = nid[ok & ~il_i]
        if root is None:
            root = int(nid[0])
        lo = np.concatenate([lo_i, mid])
        hi = np.concatenate([mid, hi_i])
        par = np.concatenate([nid, nid])
        is_left = np.concatenate([np.ones(nid.size, bool), np.zeros(nid.size, bool)])
    return _finish(parent, left, right, root, seed + 1000003, names)


def yule_tree(n_leaves, seed=1, names=False):
    """Random binary tree under the Yule (equal-rates Markov) model: a clade of m
    leaves splits into (k, m-k) with k uniform on 1..m-1, which is the
    distribution obtained by repeatedly splitting a uniformly random leaf."""
    rng = np.random.default_rng(seed)

    def split(size):
        return 1 + (rng.random(size.shape[0]) * (size - 1)).astype(np.int64).clip(0, size - 2)

    return _split_tree(int(n_leaves), split, seed, names)


def balanced_tree(n_leaves, seed=1, names=False):
    """Every clade split in halves (depth = ceil(log2 n_leaves))."""
    return _split_tree(int(n_leaves), lambda size: size // 2, seed, names)


def caterpillar_tree(n_leaves, seed=1, names=False):
    """Left comb ((((L0,L1),L2),L3),...): depth n_leaves-1, the worst case for the
    reference's O(depth^2) MRCA scan (MuchTree.pyx:1015-1028)."""
    L = int(n_leaves)
    n = 2 * L - 1
    parent = np.full(n, -1, np.int32)
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    if L > 1:
        internal = np.arange(1, n, 2, dtype=np.int64)          # ids 1,3,...,n-2
        left[internal] = internal - 2
        left[1] = 0
        right[internal] = internal + 1
        parent[internal[:-1]] = internal[:-1] + 2               # next internal up the comb
        parent[internal + 1] = internal                          # right leaves
        parent[0] = 1
        root = n - 2
    else:
        root = 0
    return _finish(parent, left, right, root, seed + 1000003, names)


def to_newick(ft, precision=None):
    """Iterative NEWICK writer (for feeding the same tree to the reference)."""
    names = {v: k for k, v in ft.leaves.items()} if ft.leaves else None
    out = []
    stack = [(ft.root, 0)]
    while stack:
        v, state = stack.pop()
        if ft.left[v] == -1:
            out.append(names[v] if names else "L%d" % (v // 2))
        elif state == 0:
            out.append("(")
            stack.append((v, 1))
            stack.append((int(ft.left[v]), 0))
            continue
        elif state == 1:
            out.append(",")
            stack.append((v, 2))
            stack.append((int(ft.right[v]), 0))
            continue
        else:
            out.append(")")
        if v != ft.root:
            # repr of the fp32 value widened to fp64 round-trips exactly through float()
            out.append(":" + repr(float(ft.distance[v])))
    out.append(";")
    return "".join(out)
