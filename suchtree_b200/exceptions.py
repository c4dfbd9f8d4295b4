"""Exception types of the drop-in API.

Same class names, hierarchy, constructor arguments, attributes and message
texts as the reference's SuchTree/exceptions.py:2-38, so that user code which
catches (or inspects) them keeps working unchanged.
"""


class SuchTreeError(Exception):
    """Root of every error raised by this package."""


class NodeNotFoundError(SuchTreeError):
    """A leaf name (or node) that the tree does not contain."""

    def __init__(self, node, message=None):
        self.node = node
        if message is None:
            message = (
                f"Leaf name not found: {node}." if isinstance(node, str) else f"Node not found: {node}"
            )
        super().__init__(message)


class InvalidNodeError(SuchTreeError):
    """A node id outside [0, size) or of the wrong kind."""

    def __init__(self, node_id, tree_size=None, message=None):
        self.node_id = node_id
        self.tree_size = tree_size
        if message is None:
            if tree_size is None:
                message = f"Invalid node ID: {node_id}"
            else:
                message = f"Node ID {node_id} out of bounds (tree size: {tree_size})"
        super().__init__(message)


class TreeStructureError(SuchTreeError):
    """The input tree is not a valid strictly-bifurcating tree."""
