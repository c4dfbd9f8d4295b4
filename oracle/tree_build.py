"""TEST INFRASTRUCTURE ONLY.

Restatement of SuchTree.__init__ (MuchTree.pyx:126-228) on top of the dendropy
stand-in (oracle/newick_ref.py): NEWICK -> the reference's Node fields as flat
arrays, with the reference's id assignment (in-order rank), epsilon
substitution and depth definition.  Independent of the product's array-based
flattener in suchtree_b200/newick.py; the two are compared node for node in
tests/, and both against the unmodified reference (tests/golden/).
"""
import os
import sys
from urllib.parse import urlparse

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import newick_ref  # noqa: E402

EPSILON = float(np.finfo(np.float64).eps)  # MuchTree.pyx:136


def load_tree(tree_input):
    """Input dispatch of MuchTree.pyx:138-155."""
    kw = dict(schema="newick", preserve_underscores=True, suppress_internal_node_taxa=True)
    if urlparse(tree_input).scheme in ("http", "https", "ftp"):
        return newick_ref.Tree.get(url=tree_input, **kw)
    if all(
        [
            "(" in tree_input,
            ")" in tree_input,
            tree_input.count("(") == tree_input.count(")"),
            tree_input.endswith(";"),
        ]
    ):
        return newick_ref.Tree.get(data=tree_input, **kw)
    return newick_ref.Tree.get(file=open(tree_input), **kw)


def build_arrays(tree_input):
    """Returns dict(parent,left,right int32; distance,support float32; leaves
    {name:id} in in-order; root; depth; size) exactly as the reference fills
    its Node array."""
    t = load_tree(tree_input)
    t.resolve_polytomies()  # :157
    size = len(t.nodes())  # :158
    parent = np.empty(size, np.int32)
    left = np.empty(size, np.int32)
    right = np.empty(size, np.int32)
    distance = np.empty(size, np.float32)
    support = np.empty(size, np.float32)
    leaves = {}
    internal = []
    # pass 1: ids + name maps (:171-180)
    for node_id, node in enumerate(t.inorder_node_iter()):
        node.node_id = node_id
        if node.taxon:
            leaves[node.taxon.label] = node_id
        else:
            internal.append(node_id)
    root = -1
    n_leaves = 0
    # pass 2: fill nodes (:182-216)
    for node_id, node in enumerate(t.inorder_node_iter()):
        if not node.parent_node:
            d, p, root = -1.0, -1, node_id
        else:
            if not node.edge_length:  # None or 0 -> epsilon (:188-192)
                d = EPSILON
            else:
                d = node.edge_length
            p = node.parent_node.node_id
        if node.taxon:
            l = r = -1
            n_leaves += 1
        else:
            lc, rc = node.child_nodes()  # unary/nullary internal -> ValueError (:200)
            l, r = lc.node_id, rc.node_id
        try:
            s = float(node.label)  # :207-210
        except (TypeError, ValueError):
            s = -1
        parent[node_id], left[node_id], right[node_id] = p, l, r
        distance[node_id], support[node_id] = d, s  # C float stores (:55-60)
    # depth (:218-225)
    depth = 0
    for node_id in leaves.values():
        n = 1
        while parent[node_id] != -1:
            node_id = parent[node_id]
            n += 1
        depth = max(depth, n)
    return dict(
        parent=parent,
        left=left,
        right=right,
        distance=distance,
        support=support,
        leaves=leaves,
        internal_nodes=np.array(internal, dtype=np.int64),
        root=root,
        depth=depth,
        size=size,
        n_leaves=n_leaves,
    )
