"""NEWICK text -> the flat node arrays the device index is built from.

Host-side replacement for the dendropy calls in SuchTree.__init__
(MuchTree.pyx:138-157, 171, 182, 200) plus its two fill passes (:171-216).  The
rules that DEFINE node ids are reproduced exactly, because every id a user holds
must stay bit-identical:

  * leaf labels are taxa, internal labels are support values; `[...]` comments
    are dropped; 'quoted labels' keep their text ('' = escaped quote); a quote
    inside a bare label is an ordinary character;
    underscores are preserved (preserve_underscores=True, :141);
  * polytomies are resolved as dendropy's deterministic resolve_polytomies()
    does (:157): nodes with >2 children are collected in post-order, then the
    first two children are repeatedly re-attached under a new zero-length node
    appended at the END of the child list, until two remain;
  * ids are in-order ranks of the binarised tree (:171-180) -- leaves get even
    ids, internal nodes odd ids;
  * a missing or zero branch length becomes the polytomy epsilon 2.22e-16
    (:136, :188-194); lengths are stored as fp32 (:55-60); the root gets the -1
    sentinel (:183-186); support = float(label) or -1 (:207-210).

flatten() runs the native loader of libsuchtree_b200.so (csrc/st_newick.cu, host
C++, O(n), iterative: ml.tree's 108,653 nodes in ~20 ms, a 10^6-deep caterpillar
loads fine).  flatten_py() is the same algorithm in Python, kept as an independent
statement of the rules that the CPU tests compare the native loader against.
"""
import re

import numpy as np

from .exceptions import TreeStructureError

EPSILON = float(np.finfo(np.float64).eps)

# a quote opens a quoted label only at the START of a token; inside a bare label it is an
# ordinary character (Cuphea_o'donellii in the reference's data/plant-pollinators/rabr)
_TOKENS = re.compile(r"\[[^\]]*\]|'(?:[^']|'')*'|[(),:;]|[^\s()\[\]',:;][^\s()\[\],:;]*")


class FlatTree:
    """parent/left/right int32, distance/support float32, in the reference's ids."""

    __slots__ = ("parent", "left", "right", "distance", "support", "leaves", "internal_nodes", "root", "size", "n_leaves")


def _tokenize(text):
    return _TOKENS.findall(text)


def parse_newick(text):
    """Returns (children lists, label list, length list) in creation order; node 0 is the root."""
    children = [[]]
    parent = [-1]
    label = [None]
    length = [None]
    cur = 0
    expect_len = False
    seen = False
    for tok in _tokenize(text):
        c = tok[0]
        if c == "[":
            continue
        seen = True
        if len(tok) == 1 and c in "(),:;":
            if c == "(":
                k = len(children)
                children.append([]); parent.append(cur); label.append(None); length.append(None)
                children[cur].append(k)
                cur = k
            elif c == ",":
                p = parent[cur]
                if p < 0:
                    raise TreeStructureError("NEWICK: ',' outside parentheses")
                k = len(children)
                children.append([]); parent.append(p); label.append(None); length.append(None)
                children[p].append(k)
                cur = k
                expect_len = False
            elif c == ")":
                cur = parent[cur]
                if cur < 0:
                    raise TreeStructureError("NEWICK: unbalanced ')'")
                expect_len = False
            elif c == ":":
                expect_len = True
            else:  # ';' ends the first tree
                break
        elif expect_len:
            try:
                length[cur] = float(tok)
            except ValueError:
                raise TreeStructureError("NEWICK: bad branch length %r" % tok)
            expect_len = False
        else:
            label[cur] = tok[1:-1].replace("''", "'") if c == "'" else tok
    if not seen:
        raise TreeStructureError("empty NEWICK input")
    if cur != 0:
        raise TreeStructureError("NEWICK: unbalanced '('")
    return children, label, length


def _resolve_polytomies(children, label, length):
    # post-order collection first (dendropy collects, then edits)
    order = []
    stack = [(0, False)]
    while stack:
        v, done = stack.pop()
        if done or not children[v]:
            if len(children[v]) > 2:
                order.append(v)
        else:
            stack.append((v, True))
            for ch in reversed(children[v]):
                stack.append((ch, False))
    for v in order:
        ch = children[v]
        while len(ch) > 2:
            k = len(children)
            children.append([ch[0], ch[1]])
            label.append(None)
            length.append(0.0)
            del ch[0:2]
            ch.append(k)


def flatten(text):
    """NEWICK text -> FlatTree with the reference's ids and field values (native loader)."""
    import ctypes as C

    from . import _lib

    L = _lib.lib()
    raw = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    h = C.c_void_p()
    rc = L.st_newick_parse(raw, len(raw), C.byref(h))
    if rc != 0:
        raise TreeStructureError(_lib.last_error())
    try:
        n, nl, root, nb = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64()
        L.st_newick_info(h, C.byref(n), C.byref(nl), C.byref(root), C.byref(nb))
        ft = FlatTree()
        ft.size, ft.n_leaves, ft.root = int(n.value), int(nl.value), int(root.value)
        ft.parent = np.empty(ft.size, np.int32)
        ft.left = np.empty(ft.size, np.int32)
        ft.right = np.empty(ft.size, np.int32)
        ft.distance = np.empty(ft.size, np.float32)
        ft.support = np.empty(ft.size, np.float32)
        L.st_newick_arrays(h, ft.parent.ctypes.data, ft.left.ctypes.data, ft.right.ctypes.data,
                           ft.distance.ctypes.data, ft.support.ctypes.data)
        ids = np.empty(ft.n_leaves, np.int32)
        offs = np.empty(ft.n_leaves + 1, np.int64)
        names = C.create_string_buffer(max(int(nb.value), 1))
        L.st_newick_leaves(h, ids.ctypes.data, offs.ctypes.data, names)
    finally:
        L.st_newick_free(h)
    blob = names.raw[: int(nb.value)]
    o = offs.tolist()
    # same insertion order (ascending id) and overwrite-on-duplicate as the reference's dict
    ft.leaves = {blob[o[k]:o[k + 1]].decode("utf-8", "replace"): i for k, i in enumerate(ids.tolist())}
    ft.internal_nodes = np.nonzero(ft.left != -1)[0].astype(np.int64)
    return ft


def flatten_py(text):
    """The same rules in pure Python (independent restatement for the tests)."""
    children, label, length = parse_newick(text)
    _resolve_polytomies(children, label, length)
    n = len(children)
    for v in range(n):
        k = len(children[v])
        if k == 1:
            raise TreeStructureError(
                "node with a single child: SuchTree requires a strictly bifurcating tree")
        if k == 0 and label[v] is None:
            raise TreeStructureError("leaf without a name")
    # in-order ranks, iteratively
    new_id = np.empty(n, np.int64)
    order = []
    stack = [(0, False)]
    while stack:
        v, emit = stack.pop()
        ch = children[v]
        if emit or not ch:
            new_id[v] = len(order)
            order.append(v)
        else:
            stack.append((ch[1], False))
            stack.append((v, True))
            stack.append((ch[0], False))
    ft = FlatTree()
    ft.size = n
    ft.parent = np.full(n, -1, np.int32)
    ft.left = np.full(n, -1, np.int32)
    ft.right = np.full(n, -1, np.int32)
    dist = np.empty(n, np.float64)
    ft.support = np.full(n, -1.0, np.float32)
    ft.leaves = {}
    internal = []
    ft.n_leaves = 0
    for i, v in enumerate(order):
        ch = children[v]
        if ch:
            l, r = int(new_id[ch[0]]), int(new_id[ch[1]])
            ft.left[i], ft.right[i] = l, r
            ft.parent[l] = i
            ft.parent[r] = i
            internal.append(i)
            if label[v] is not None:
                try:
                    ft.support[i] = float(label[v])
                except ValueError:
                    pass
        else:
            ft.leaves[label[v]] = i
            ft.n_leaves += 1
        ln = length[v]
        dist[i] = ln if ln else EPSILON  # None, 0.0 and -0.0 -> epsilon
    ft.root = int(new_id[0])
    dist[ft.root] = -1.0
    with np.errstate(over="ignore"):
        ft.distance = dist.astype(np.float32)
    ft.internal_nodes = np.array(internal, dtype=np.int64)
    return ft
