"""The rest of SuchTree's Python surface: node queries, traversals, topology helpers and
graph / matrix exports.

None of this is on the accelerated path -- these are host-side walks over the flattened
node arrays (parent / left / right / distance / support, the reference's `Node` fields,
MuchTree.pyx:55-60) that the reference implements as Python loops over the same fields.
They are here so that a caller of the reference finds every method under the same name,
with the same argument forms, orders, return types and exceptions.  Where a batched form
exists on the device (MRCA, distances, clade intervals) the method uses it.

Each method cites the reference lines it mirrors; golden vectors produced by the
unmodified reference pin them (tests/golden/make_golden_api.py -> api.json).
"""
from collections import deque

import numpy as np

from .exceptions import InvalidNodeError, NodeNotFoundError  # noqa: F401  (re-exported for callers)


def _f(x):
    """float32 node field -> Python float, as Cython hands a C float to Python."""
    return float(x)


class TreeExtras:
    """Mixin for SuchTree.  Uses self._ft (FlatTree), self._root, self._size, self.leaves,
    self.leaf_nodes, self._validate_node(), and the accelerated entry points."""

    # ====== node queries (MuchTree.pyx:370-612) ======
    def get_ancestors(self, node):
        """Generator over ancestor ids from the parent up to the root; MuchTree.pyx:370-394."""
        node_id = self._validate_node(node)
        parent = self._ft.parent
        while True:
            p = int(parent[node_id])
            if p == -1:
                break
            yield p
            node_id = p

    def get_support(self, node):
        """Support value of a node (-1 when there is none); MuchTree.pyx:465-481."""
        return _f(self._ft.support[self._validate_node(node)])

    def _bfs_nodes(self, start):
        """ids below (and including) `start` in the reference's queue order: a list that grows
        while it is walked, left child before right child (MuchTree.pyx:589-612)."""
        left, right = self._ft.left, self._ft.right
        cur = np.array([start], dtype=np.int64)
        levels = []
        while cur.size:
            levels.append(cur)
            inner = cur[left[cur] != -1]
            nxt = np.empty(2 * inner.size, dtype=np.int64)
            nxt[0::2] = left[inner]
            nxt[1::2] = right[inner]
            cur = nxt
        return np.concatenate(levels)

    def get_nodes(self, from_node=-1):
        """All node ids below from_node (default: the root) in breadth-first queue order.
        The reference defines this twice; the second definition (from_node=-1, no
        validation) is the live one: MuchTree.pyx:589-612."""
        start = self._root if from_node == -1 else int(from_node)
        return self._bfs_nodes(start)

    def get_internal_nodes(self, from_node=-1):
        """Internal node ids below from_node in breadth-first queue order; MuchTree.pyx:560-587
        (the live, second definition)."""
        start = self._root if from_node == -1 else int(from_node)
        ids = self._bfs_nodes(start)
        return ids[self._ft.left[ids] != -1]

    # ====== node tests (MuchTree.pyx:681-811) ======
    def is_descendant(self, descendant, ancestor):
        """True when `descendant` lies below `ancestor`; MuchTree.pyx:681-701."""
        a, d = self._validate_node_pair(ancestor, descendant)
        return self.is_ancestor(a, d) == 1

    def is_sibling(self, node1, node2):
        """MuchTree.pyx:749-775."""
        a, b = self._validate_node_pair(node1, node2)
        if a == self._root or b == self._root:
            return False
        pa, pb = int(self._ft.parent[a]), int(self._ft.parent[b])
        return pa == pb and pa != -1

    def has_children(self, node):
        return self.is_internal(node)

    def has_parent(self, node):
        return not self.is_root(node)

    # ====== validation helpers (MuchTree.pyx:2302-2370) ======
    def _validate_leaf_node(self, node):
        node_id = self._validate_node(node)
        if self._ft.left[node_id] != -1:
            raise InvalidNodeError(node_id, message="Node {node_id} is not a leaf node".format(node_id=str(node_id)))
        return node_id

    def _validate_internal_node(self, node):
        node_id = self._validate_node(node)
        if self._ft.left[node_id] == -1:
            raise InvalidNodeError(node_id, message="Node {node_id} is not an internal node".format(node_id=str(node_id)))
        return node_id

    def _convert_to_leaf_names(self, node_ids):
        names = []
        leaf_nodes = self.leaf_nodes
        left = self._ft.left
        for node_id in node_ids:
            if left[node_id] != -1:
                raise InvalidNodeError(node_id, message="Node {node_id} is not a leaf".format(node_id=str(node_id)))
            names.append(leaf_nodes[int(node_id)])
        return names

    # ====== topology (MuchTree.pyx:1151-1200, 1424-1463) ======
    def bipartition(self, node, by_id=False):
        """The two leaf sets split by an internal node; MuchTree.pyx:1151-1184."""
        node_id = self._validate_internal_node(node)
        l, r = int(self._ft.left[node_id]), int(self._ft.right[node_id])
        ll, rl = self.get_leaves(l), self.get_leaves(r)
        if by_id:
            return frozenset((frozenset(ll.tolist()), frozenset(rl.tolist())))
        return frozenset((frozenset(self._convert_to_leaf_names(ll)), frozenset(self._convert_to_leaf_names(rl))))

    def bipartitions(self, by_id=False):
        """Generator over the bipartitions of every internal node, in get_internal_nodes()
        order; MuchTree.pyx:1186-1200."""
        for node_id in self.get_internal_nodes():
            yield self.bipartition(int(node_id), by_id=by_id)

    def path_between_nodes(self, a, b):
        """Node ids on the path a -> MRCA -> b; MuchTree.pyx:1424-1463."""
        node_a, node_b = self._validate_node_pair(a, b)
        if node_a == node_b:
            return [node_a]
        mrca = self.common_ancestor(node_a, node_b)
        parent = self._ft.parent
        path_a, cur = [], node_a
        while cur != mrca:
            path_a.append(cur)
            cur = int(parent[cur])
        path_b, cur = [], node_b
        while cur != mrca:
            path_b.append(cur)
            cur = int(parent[cur])
        return path_a + [mrca] + list(reversed(path_b))

    # ====== traversals (MuchTree.pyx:1467-1746) ======
    def _start(self, from_node):
        return self._root if from_node is None else self._validate_node(from_node)

    def traverse_inorder(self, include_distances=True):
        """In-order walk from the root: ids, or (id, distance to parent); MuchTree.pyx:1467-1500."""
        left, right, dist = self._ft.left, self._ft.right, self._ft.distance
        cur, stack = self._root, []
        while True:
            if cur != -1:
                stack.append(cur)
                cur = int(left[cur])
            elif stack:
                cur = stack.pop()
                yield (cur, _f(dist[cur])) if include_distances else cur
                cur = int(right[cur])
            else:
                break

    def traverse_postorder(self, from_node=None):
        """MuchTree.pyx:1538-1576."""
        left, right = self._ft.left, self._ft.right
        stack, last, cur = [], None, self._start(from_node)
        while stack or cur != -1:
            if cur != -1:
                stack.append(cur)
                cur = int(left[cur])
            else:
                peek = stack[-1]
                r = int(right[peek])
                if r != -1 and last != r:
                    cur = r
                else:
                    yield peek
                    last = stack.pop()

    def traverse_levelorder(self, from_node=None):
        """MuchTree.pyx:1578-1612."""
        left, right = self._ft.left, self._ft.right
        queue = deque([self._start(from_node)])
        while queue:
            cur = queue.popleft()
            yield cur
            l, r = int(left[cur]), int(right[cur])
            if l != -1:
                queue.append(l)
            if r != -1:
                queue.append(r)

    def traverse_leaves_only(self, from_node=None):
        """Leaves in preorder; MuchTree.pyx:1614-1638."""
        left = self._ft.left
        for node_id in self.traverse_preorder(self._start(from_node)):
            if left[node_id] == -1:
                yield node_id

    def traverse_internal_only(self, from_node=None):
        """Internal nodes in preorder; MuchTree.pyx:1640-1664."""
        left = self._ft.left
        for node_id in self.traverse_preorder(self._start(from_node)):
            if left[node_id] != -1:
                yield node_id

    def traverse_with_depth(self, from_node=None):
        """(id, depth below the start node) in preorder; MuchTree.pyx:1666-1701."""
        left, right = self._ft.left, self._ft.right
        stack = [(self._start(from_node), 0)]
        while stack:
            cur, depth = stack.pop()
            yield (cur, depth)
            r, l = int(right[cur]), int(left[cur])
            if r != -1:
                stack.append((r, depth + 1))
            if l != -1:
                stack.append((l, depth + 1))

    def traverse_with_distances(self, from_node=None):
        """(id, distance to parent, cumulative distance) in preorder.  As in the reference
        (MuchTree.pyx:1703-1746) the cumulative figure of a node is the sum of the edges ABOVE
        it below the start node's own edge, i.e. it excludes the node's own edge, and the start
        node's edge is counted unless it is the root's -1 sentinel."""
        left, right, dist = self._ft.left, self._ft.right, self._ft.distance
        stack = [(self._start(from_node), 0.0)]
        while stack:
            cur, to_root = stack.pop()
            d = _f(dist[cur])
            yield (cur, d, to_root)
            nxt = to_root + (d if d != -1 else 0)
            r, l = int(right[cur]), int(left[cur])
            if r != -1:
                stack.append((r, nxt))
            if l != -1:
                stack.append((l, nxt))

    def in_order(self, distances=True):
        from .tree import _deprecation_warning

        _deprecation_warning("in_order()", "traverse_inorder()")
        return self.traverse_inorder(include_distances=distances)

    # ====== graph / matrix exports (MuchTree.pyx:1750-1989) ======
    def adjacency_matrix(self, from_node=None):
        """Weighted adjacency matrix of the clade below from_node, nodes in breadth-first
        queue order; zero-length edges are replaced by polytomy_epsilon; MuchTree.pyx:1750-1815."""
        node_ids = self._bfs_nodes(self._start(from_node))
        n = node_ids.shape[0]
        adj = np.zeros((n, n), dtype=float)
        index_of = np.full(self._size, -1, dtype=np.int64)
        index_of[node_ids] = np.arange(n)
        parent = np.asarray(self._ft.parent, dtype=np.int64)[node_ids]
        d = np.asarray(self._ft.distance, dtype=np.float64)[node_ids]
        d = np.where(d == 0, d + self.polytomy_epsilon, d)
        pidx = np.where(parent >= 0, index_of[np.maximum(parent, 0)], -1)
        rows = np.nonzero(pidx >= 0)[0]  # the start node's parent is outside the clade (or -1)
        adj[rows, pidx[rows]] = d[rows]
        adj[pidx[rows], rows] = d[rows]
        return {"adjacency_matrix": adj, "node_ids": node_ids}

    def laplacian_matrix(self, from_node=None):
        """MuchTree.pyx:1817-1854."""
        res = self.adjacency_matrix(self._start(from_node))
        adj = res["adjacency_matrix"]
        lap = np.zeros(adj.shape, dtype=float)
        np.fill_diagonal(lap, adj.sum(axis=0))
        return {"laplacian": lap - adj, "node_ids": res["node_ids"]}

    def incidence_matrix(self, from_node=None):
        """Node-by-edge incidence (+1 parent end, -1 child end) and the edge list; the start
        node's own edge to ITS parent is listed too when it has one, as in the reference, where
        looking that parent up then fails (MuchTree.pyx:1856-1917)."""
        node_ids = self._bfs_nodes(self._start(from_node))
        parent = self._ft.parent
        edges = [(int(parent[c]), int(c)) for c in node_ids if parent[c] != -1]
        index_of = {int(v): i for i, v in enumerate(node_ids)}
        inc = np.zeros((node_ids.shape[0], len(edges)), dtype=int)
        for e, (p, c) in enumerate(edges):
            if p not in index_of:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # np.where(...)[0][0] upstream
            inc[index_of[p], e] = 1
            inc[index_of[c], e] = -1
        return {"incidence_matrix": inc, "node_ids": node_ids, "edge_list": edges}

    def degree_sequence(self, from_node=None):
        """MuchTree.pyx:1958-1989."""
        res = self.adjacency_matrix(from_node)
        degrees = np.sum(res["adjacency_matrix"] > 0, axis=1)
        return {"degrees": degrees, "node_ids": res["node_ids"], "max_degree": degrees.max(),
                "min_degree": degrees.min()}

    def adjacency(self, node=-1):
        from .tree import _deprecation_warning

        _deprecation_warning("adjacency()", "adjacency_matrix()")
        return self.adjacency_matrix(None if node == -1 else node)

    def laplacian(self, node=-1):
        from .tree import _deprecation_warning

        _deprecation_warning("laplacian()", "laplacian_matrix()")
        return self.laplacian_matrix(None if node == -1 else node)

    # ====== links to a SuchLinkedTrees column (MuchTree.pyx:1993-2014) ======
    # The reference overwrites a linked leaf's right_child with the column id; here the
    # node arrays stay immutable (the device index is built from them) and the column ids
    # live in a side table.
    def link_leaf(self, leaf_id, col_id):
        if self._ft.left[leaf_id] != -1:
            raise Exception("Cannot link non-leaf node.", leaf_id)
        if leaf_id not in self.leaf_nodes:
            raise Exception("Unknown leaf id.", leaf_id)
        if getattr(self, "_leaf_links", None) is None:
            self._leaf_links = {}
        self._leaf_links[int(leaf_id)] = int(col_id)

    def get_links(self, leaf_ids):
        if not set(leaf_ids) <= set(self.leaf_nodes.keys()):
            raise Exception("Unknown leaf id(s).", leaf_ids)
        links = getattr(self, "_leaf_links", None) or {}
        return np.array([links.get(int(leaf), -1) for leaf in leaf_ids], dtype=int)

    # ====== exports (MuchTree.pyx:2018-2251) ======
    def _depth_below_root(self, node_id):
        depth, cur, parent = 0, node_id, self._ft.parent
        while cur != self._root and parent[cur] != -1:
            cur = int(parent[cur])
            depth += 1
        return depth

    def to_networkx_nodes(self, from_node=None):
        """(node_id, attributes) in get_descendants() order; MuchTree.pyx:2018-2075."""
        ft = self._ft
        for node_id in self.get_descendants(self._start(from_node)):
            attributes = {}
            if ft.left[node_id] == -1:
                attributes["type"] = "leaf"
                attributes["label"] = self.leaf_nodes[node_id]
            else:
                attributes["type"] = "internal"
                attributes["label"] = "node_%d" % node_id
            support = _f(ft.support[node_id])
            if support != -1:
                attributes["support"] = support
            distance = _f(ft.distance[node_id])
            if distance != -1:
                attributes["distance_to_parent"] = distance
            attributes["distance_to_root"] = self.distance_to_root(node_id)
            attributes["depth"] = self._depth_below_root(node_id)
            yield (node_id, attributes)

    def to_networkx_edges(self, from_node=None):
        """(child, parent, attributes) in get_descendants() order; MuchTree.pyx:2077-2123."""
        ft = self._ft
        for node_id in self.get_descendants(self._start(from_node)):
            parent_id = int(ft.parent[node_id])
            if parent_id == -1:
                continue
            d = _f(ft.distance[node_id])
            attributes = {"weight": d, "length": d}
            if ft.left[node_id] != -1:
                support = _f(ft.support[node_id])
                if support != -1:
                    attributes["support"] = support
            yield (node_id, parent_id, attributes)

    def to_networkx_graph(self, from_node=None):
        """MuchTree.pyx:2125-2156."""
        try:
            import networkx as nx
        except ImportError:
            raise ImportError("NetworkX is required for to_networkx_graph()")
        G = nx.Graph()
        for node_id, attributes in self.to_networkx_nodes(from_node):
            G.add_node(node_id, **attributes)
        for child_id, parent_id, attributes in self.to_networkx_edges(from_node):
            G.add_edge(child_id, parent_id, **attributes)
        return G

    def nodes_data(self):
        from .tree import _deprecation_warning

        _deprecation_warning("nodes_data()", "to_networkx_nodes()")
        return self.to_networkx_nodes()

    def edges_data(self):
        from .tree import _deprecation_warning

        _deprecation_warning("edges_data()", "to_networkx_edges()")
        return self.to_networkx_edges()

    def to_newick(self, from_node=None, include_support=True, include_distances=True):
        """NEWICK text of the clade below from_node; MuchTree.pyx:2181-2229.  Same text as the
        reference's recursion (str() of the float32 fields widened to double), built with an
        explicit stack so that 10^6-deep ladders do not hit the recursion limit."""
        start = self._start(from_node)
        ft, names = self._ft, self.leaf_nodes
        out, stack = [], [(start, 0)]
        while stack:
            node_id, state = stack.pop()
            l, r = int(ft.left[node_id]), int(ft.right[node_id])
            if l == -1:
                out.append(names[node_id])
            elif state == 0:
                out.append("(")
                stack.append((node_id, 1))
                stack.append((l, 0))
                continue
            elif state == 1:
                out.append(",")
                stack.append((node_id, 2))
                stack.append((r, 0))
                continue
            else:
                out.append(")")
                if include_support:
                    support = _f(ft.support[node_id])
                    if support != -1:
                        out.append(str(support))
            if include_distances and node_id != start:
                distance = _f(ft.distance[node_id])
                if distance != -1:
                    out.append(":%s" % distance)
        return "".join(out) + ";"

    def dump_array(self):
        """Print the node array; MuchTree.pyx:2231-2240."""
        ft = self._ft
        for n in range(self._size):
            print("id : %d ->" % n)
            print("   distance    : %0.3f" % ft.distance[n])
            print("   parent      : %d" % ft.parent[n])
            print("   left child  : %d" % ft.left[n])
            print("   right child : %d" % ft.right[n])

    # ====== deprecated aliases (MuchTree.pyx:2418-2470) ======
    def get_lineage(self, node):
        from .tree import _deprecation_warning

        _deprecation_warning("get_lineage()", "get_ancestors()")
        return self.get_ancestors(node)

    def get_descendant_nodes(self, node):
        from .tree import _deprecation_warning

        _deprecation_warning("get_descendant_nodes()", "get_descendants()")
        return self.get_descendants(node)

    def get_leafs(self, node):
        from .tree import _deprecation_warning

        _deprecation_warning("get_leafs()", "get_leaves()")
        return self.get_leaves(node)

    def is_internal_node(self, node):
        from .tree import _deprecation_warning

        _deprecation_warning("is_internal_node()", "is_internal()")
        return self.is_internal(node)

    def get_bipartition(self, node, by_id=False):
        from .tree import _deprecation_warning

        _deprecation_warning("get_bipartition()", "bipartition()")
        return self.bipartition(node, by_id=by_id)


class LinkedExtras:
    """Mixin for SuchLinkedTrees: the joint graph of the two trees and their links
    (MuchTree.pyx:3081-3208)."""

    def adjacency(self, deletions=0, additions=0, swaps=0):
        """Adjacency matrix of TreeA's clade, TreeB's clade and the links between them, tree
        blocks scaled to a maximum of 1, links weighted by the mean of the two trees' mean scaled
        edge; MuchTree.pyx:3081-3131 (including its `range(1, k)` perturbation counts)."""
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            TA = self.TreeA.adjacency(node=self.subset_a_root)
            TB = self.TreeB.adjacency(node=self.subset_b_root)
            eps_a, eps_b = self.TreeA.polytomy_distance, self.TreeB.polytomy_distance
        ta_aj, tb_aj = TA["adjacency_matrix"], TB["adjacency_matrix"]
        ll = np.array(self.linklist)
        for _ in range(1, deletions):
            ll = np.delete(ll, np.random.randint(len(ll)), axis=0)
        for _ in range(1, swaps):
            x, y = np.random.choice(range(len(ll)), size=2, replace=False)
            X, Y = ll[x, 1], ll[y, 1]
            ll[x, 1] = Y
            ll[y, 1] = X
        for _ in range(1, additions):
            a = np.random.choice(list(self.TreeA.leaves.values()))
            b = np.random.choice(list(self.TreeB.leaves.values()))
            ll = np.concatenate((ll, np.array([[b, a]])), axis=0)
        na, nb = ta_aj.shape[0], tb_aj.shape[0]
        pos_a = np.full(self.TreeA.size, -1, dtype=np.int64)
        pos_a[TA["node_ids"]] = np.arange(na)
        pos_b = np.full(self.TreeB.size, -1, dtype=np.int64)
        pos_b[TB["node_ids"]] = np.arange(nb)
        ta_links, tb_links = pos_a[ll[:, 1]], pos_b[ll[:, 0]]
        if (ta_links < 0).any() or (tb_links < 0).any():
            bad = ll[np.argmax((ta_links < 0) | (tb_links < 0))]
            raise ValueError("%d is not in list" % (bad[1] if pos_a[bad[1]] < 0 else bad[0]))  # list.index upstream
        aj = np.zeros((na + nb, na + nb))
        aj[:na, :na] = ta_aj / ta_aj.max()
        aj[na:, na:] = tb_aj / tb_aj.max()
        ta_mean = np.mean(ta_aj.flatten()[ta_aj.flatten() > eps_a])
        tb_mean = np.mean(tb_aj.flatten()[tb_aj.flatten() > eps_b])
        link_mean = (ta_mean / ta_aj.max() + tb_mean / tb_aj.max()) / 2.0
        aj[tb_links + na, ta_links] = link_mean
        aj[ta_links, tb_links + na] = link_mean
        return aj

    def laplacian(self, deletions=0, additions=0, swaps=0):
        """MuchTree.pyx:3133-3145."""
        aj = self.adjacency(deletions=deletions, additions=additions, swaps=swaps)
        lp = np.zeros(aj.shape)
        np.fill_diagonal(lp, aj.sum(axis=0))
        return lp - aj

    def spectrum(self, deletions=0, additions=0, swaps=0):
        """Eigenvalues of the graph Laplacian in ascending order (the reference calls LAPACK
        dsyev('N','U'), MuchTree.pyx:3147-3173; numpy's eigvalsh is the same family of solver)."""
        return np.linalg.eigvalsh(self.laplacian(deletions, additions, swaps), UPLO="U")

    def to_igraph(self, deletions=0, additions=0, swaps=0):
        """MuchTree.pyx:3175-3197."""
        try:
            from igraph import ADJ_UNDIRECTED, Graph
        except ImportError:
            raise Exception("igraph package not installed.")
        g = Graph.Weighted_Adjacency(self.adjacency(deletions=deletions, additions=additions, swaps=swaps).tolist(),
                                     mode=ADJ_UNDIRECTED)
        na = len(list(self.TreeA.get_descendants(self.subset_a_root)))
        nb = len(list(self.TreeB.get_descendants(self.subset_b_root)))
        g.vs["color"] = ["#e1e329ff"] * na + ["#24878dff"] * nb
        g.vs["label"] = ["h" + str(i) for i in range(na)] + ["g" + str(i) for i in range(nb)]
        g.vs["tree"] = [0] * na + [1] * nb
        return g

    def dump_table(self):
        """Print the link table column by column; MuchTree.pyx:3200-3208."""
        for i in range(self.n_cols):
            col = self.get_column_leafs(i)
            print("column", i, ":", ",".join(map(str, col.tolist())))
