"""TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

Object-based NEWICK reader that stands in for the third-party `dendropy`
package, which the reference imports (MuchTree.pyx:3) but which is neither
vendored under /root/reference nor installable here (un-pinned in
requirements.txt:4, pyproject.toml:8,27).

Two users:
  * oracle/ref_loader.py registers this module as `dendropy` so that the
    UNMODIFIED reference extension (compiled from /root/reference by
    oracle/build_ref.py) can be imported and run as the parity oracle O1.
  * oracle/tree_build.py restates SuchTree.__init__ (MuchTree.pyx:126-228)
    on top of it (oracle O2).

What is restated here is dendropy's *published behaviour* at the reference's
call sites (MuchTree.pyx:139-157, 171, 182, 200):
  - Tree.get(data=|file=|url=, schema='newick', preserve_underscores=True,
    suppress_internal_node_taxa=True): leaf labels become taxa, internal
    labels stay plain node labels, [..] comments are dropped, 'quoted'
    labels keep their content ('' is an escaped quote), a missing branch
    length is None.
  - Tree.resolve_polytomies(): deterministic mode (rng=None) -- polytomies
    are collected in post-order, then for each one the first two children
    are repeatedly detached and re-attached under a new zero-length node
    that is appended at the END of the child list, until two remain.
  - Tree.inorder_node_iter(): left subtree, node, right subtree; only
    defined for strictly binary trees (TypeError otherwise).
  - Tree.nodes(): pre-order list.
Pinned against the reference's own published vectors in
tests/test_oracle_golden.py (SLT.linklist of Gopher-Louse.ipynb:1238-1263,
leaves of docs/examples/SuchTree_examples.md:99-112, test.matrix).

Everything is iterative: a 10^6-deep caterpillar must not hit the Python
recursion limit.
"""
import re

_TOKEN = re.compile(
    r"""\s+                      # whitespace
      | \[[^\]]*\]               # comment
      | '(?:[^']|'')*'           # quoted label
      | [(),:;]                  # punctuation
      | [^\s()\[\]',:;][^\s()\[\],:;]*   # bare label / number; a quote INSIDE a bare label is an
                                 # ordinary character (data/plant-pollinators/rabr/plant.tree has
                                 # the leaf Cuphea_o'donellii, and rabr_links.csv, made with real
                                 # dendropy, calls it exactly that)
    """,
    re.X,
)


class Taxon:
    __slots__ = ("label",)

    def __init__(self, label):
        self.label = label

    def __repr__(self):
        return "<Taxon %r>" % (self.label,)


class Node:
    __slots__ = ("_child_nodes", "parent_node", "edge_length", "taxon", "label", "node_id")

    def __init__(self):
        self._child_nodes = []
        self.parent_node = None
        self.edge_length = None
        self.taxon = None
        self.label = None
        self.node_id = None

    # -- the slice of dendropy.Node the reference touches --
    def child_nodes(self):
        return list(self._child_nodes)

    def add_child(self, ch):
        ch.parent_node = self
        self._child_nodes.append(ch)
        return ch

    def remove_child(self, ch):
        self._child_nodes.remove(ch)
        ch.parent_node = None
        return ch

    def is_leaf(self):
        return not self._child_nodes

    def distance_from_root(self):
        d, n = 0.0, self
        while n.parent_node is not None:
            if n.edge_length is not None:
                d += n.edge_length
            n = n.parent_node
        return d


class Tree:
    def __init__(self, seed_node=None):
        self.seed_node = seed_node if seed_node is not None else Node()

    # ---- construction -------------------------------------------------
    @classmethod
    def get(cls, **kw):
        schema = kw.get("schema", "newick")
        if schema != "newick":
            raise ValueError("stand-in supports schema='newick' only")
        if "data" in kw:
            text = kw["data"]
        elif "file" in kw:
            f = kw["file"]
            text = f.read()
            try:
                f.close()
            except Exception:
                pass
        elif "path" in kw:
            with open(kw["path"]) as f:
                text = f.read()
        elif "url" in kw:
            from urllib.request import urlopen

            text = urlopen(kw["url"]).read().decode()
        else:
            raise TypeError("need one of data=, file=, path=, url=")
        preserve = kw.get("preserve_underscores", False)
        return cls._parse(text, preserve_underscores=preserve)

    @classmethod
    def _parse(cls, text, preserve_underscores=True):
        root = Node()
        cur = root
        expecting_length = False
        seen_any = False
        pos, n = 0, len(text)
        match = _TOKEN.match
        while pos < n:
            m = match(text, pos)
            if m is None:
                raise ValueError("NEWICK syntax error at offset %d" % pos)
            tok = m.group()
            pos = m.end()
            c = tok[0]
            if c.isspace() or c == "[":
                continue
            seen_any = True
            if c == "(" and len(tok) == 1:
                cur = cur.add_child(Node())
            elif c == "," and len(tok) == 1:
                if cur.parent_node is None:
                    raise ValueError("NEWICK: ',' outside parentheses")
                cur = cur.parent_node.add_child(Node())
                expecting_length = False
            elif c == ")" and len(tok) == 1:
                if cur.parent_node is None:
                    raise ValueError("NEWICK: unbalanced ')'")
                cur = cur.parent_node
                expecting_length = False
            elif c == ":" and len(tok) == 1:
                expecting_length = True
            elif c == ";" and len(tok) == 1:
                break
            else:
                if expecting_length:
                    cur.edge_length = float(tok)
                    expecting_length = False
                else:
                    if c == "'":
                        lab = tok[1:-1].replace("''", "'")
                    else:
                        lab = tok if preserve_underscores else tok.replace("_", " ")
                    cur.label = lab
        if not seen_any:
            raise ValueError("empty NEWICK input")
        if cur is not root:
            raise ValueError("NEWICK: unbalanced '('")
        # suppress_internal_node_taxa=True: only leaves get taxa
        stack = [root]
        while stack:
            nd = stack.pop()
            if nd._child_nodes:
                stack.extend(nd._child_nodes)
            elif nd.label is not None:
                nd.taxon = Taxon(nd.label)
                nd.label = None
        return cls(root)

    # ---- traversals ---------------------------------------------------
    def preorder_node_iter(self):
        stack = [self.seed_node]
        while stack:
            nd = stack.pop()
            yield nd
            stack.extend(reversed(nd._child_nodes))

    def postorder_node_iter(self):
        stack = [(self.seed_node, False)]
        while stack:
            nd, done = stack.pop()
            if done or not nd._child_nodes:
                yield nd
            else:
                stack.append((nd, True))
                for ch in reversed(nd._child_nodes):
                    stack.append((ch, False))

    def inorder_node_iter(self):
        stack = [(self.seed_node, False)]
        while stack:
            nd, emit = stack.pop()
            k = len(nd._child_nodes)
            if emit or k == 0:
                yield nd
            elif k == 2:
                stack.append((nd._child_nodes[1], False))
                stack.append((nd, True))
                stack.append((nd._child_nodes[0], False))
            else:
                raise TypeError("In-order traversal only supported for binary trees")

    def leaf_node_iter(self):
        for nd in self.preorder_node_iter():
            if not nd._child_nodes:
                yield nd

    def nodes(self):
        return list(self.preorder_node_iter())

    # ---- polytomies ---------------------------------------------------
    def resolve_polytomies(self, limit=2):
        polytomies = [nd for nd in self.postorder_node_iter() if len(nd._child_nodes) > limit]
        for nd in polytomies:
            while len(nd._child_nodes) > limit:
                nn = Node()
                nn.edge_length = 0
                c1 = nd._child_nodes[0]
                c2 = nd._child_nodes[1]
                nd.remove_child(c1)
                nd.remove_child(c2)
                nn.add_child(c1)
                nn.add_child(c2)
                nd.add_child(nn)
