"""suchtree_b200 -- B200-native batched patristic distances, drop-in for the hot
path of ryneches/SuchTree (see SURVEY.md §8 and DESIGN.md).

    from suchtree_b200 import SuchTree, SuchLinkedTrees, pearson

Everything numerical runs in libsuchtree_b200.so (hand-written sm_100a CUDA behind
the C ABI of include/suchtree_b200.h); there is no CPU fallback.
"""
from .exceptions import InvalidNodeError, NodeNotFoundError, SuchTreeError, TreeStructureError
from .linked import SuchLinkedTrees, as_moments, moments_pearson, pearson
from .tree import SuchTree

__version__ = "0.1.0"

__all__ = [
    "SuchTree",
    "SuchLinkedTrees",
    "pearson",
    "moments_pearson",
    "as_moments",
    "SuchTreeError",
    "NodeNotFoundError",
    "InvalidNodeError",
    "TreeStructureError",
    "__version__",
]
