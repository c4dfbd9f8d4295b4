"""ctypes binding of libsuchtree_b200.so (the C ABI in include/suchtree_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc
(suchtree_b200/build.py); if that fails, or no CUDA device is visible when a
tree is created, the error propagates.
"""
import ctypes as C
import os
import threading

from . import build as _build
from .exceptions import InvalidNodeError, SuchTreeError, TreeStructureError

ST_OK = 0
ST_ERR_INVALID_ARG = 1
ST_ERR_CUDA = 2
ST_ERR_NOT_BINARY = 3
ST_ERR_NOT_INORDER = 4
ST_ERR_NODE_RANGE = 5
ST_ERR_LENGTH_MISMATCH = 6
ST_ERR_NOMEM = 7


class TreeInfo(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64),
        ("n_leaves", C.c_int64),
        ("root", C.c_int32),
        ("depth", C.c_int32),
        ("device", C.c_int32),
        ("block_shift", C.c_int32),
        ("micro_shift", C.c_int32),
        ("n_blocks", C.c_int32),
        ("index_bytes", C.c_int64),
        ("query_smem_bytes", C.c_int32),
        ("sm_count", C.c_int32),
        ("layout", C.c_int32),
    ]


class Moments(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("n", "x0", "y0", "sx", "sy", "sxx", "syy", "sxy")]


_lib = None
_lock = threading.Lock()

# name -> (restype, argtypes); every symbol include/suchtree_b200.h declares
_vp, _i64, _i32, _u64, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_int
SIGNATURES = {
    "st_last_error": (C.c_char_p, []),
    "st_version": (_int, []),
    "st_device_count": (_int, [C.POINTER(_int)]),
    "st_tree_create": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _int, _int, C.POINTER(_vp)]),
    "st_tree_create_ex": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _int, _int, _int, C.POINTER(_vp)]),
    "st_tree_destroy": (None, [_vp]),
    "st_tree_get_info": (_int, [_vp, C.POINTER(TreeInfo)]),
    "st_tree_export": (_int, [_vp, _vp, _vp, _vp]),
    "st_newick_parse": (_int, [C.c_char_p, _i64, C.POINTER(_vp)]),
    "st_newick_free": (None, [_vp]),
    "st_newick_info": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i64)]),
    "st_newick_arrays": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "st_newick_leaves": (_int, [_vp, _vp, _vp, _vp]),
    "st_bad_node": (_i64, []),
    "st_distances": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "st_mrca": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "st_distances_device": (_int, [_vp, _vp, _int, _i64, _vp, _vp, _vp]),
    "st_check_range": (_int, [_vp, _vp]),
    "st_quartet_topologies": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "st_quartet_topologies_device": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "st_random_leaf_pairs_device": (_int, [_vp, _u64, _i64, _i64, _vp, _int, _vp]),
    "st_distance_matrix": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "st_linked_distances": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "st_sample_linked_cycle": (
        _int, [_vp, _vp, _vp, _i64, C.POINTER(_u64), _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "st_sample_moments": (
        _int, [_vp, _vp, _vp, _i64, _u64, _i64, _i64, C.c_double, C.c_double, C.POINTER(Moments)]),
    "st_linked_moments": (
        _int, [_vp, _vp, _vp, _i64, _i64, _i64, C.c_double, C.c_double, C.POINTER(Moments)]),
    "st_clade_moments": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "st_moments_pearson": (C.c_double, [C.POINTER(Moments)]),
    "st_pearson": (_int, [_int, _vp, _vp, _i64, C.POINTER(C.c_double)]),
    "st_bench_pack": (_int, [_i64, _int, C.POINTER(C.c_double)]),
    "st_bench_gather": (_int, [_int, _i64, _i64, _int, C.POINTER(C.c_double)]),
}


def lib():
    """Load (building first if needed) the CUDA library.  Raises if impossible."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.LIB
                if not os.path.exists(path) or (
                    os.environ.get("SUCHTREE_B200_REBUILD") and _build.needs_build()
                ):
                    path = _build.build()
                L = C.CDLL(path)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def last_error():
    return lib().st_last_error().decode(errors="replace")


def check(rc, tree_size=None):
    """Map an st_status to the reference's exception types."""
    if rc == ST_OK:
        return
    msg = last_error()
    if rc == ST_ERR_NODE_RANGE:
        raise InvalidNodeError(int(lib().st_bad_node()), tree_size)
    if rc in (ST_ERR_NOT_BINARY, ST_ERR_NOT_INORDER):
        raise TreeStructureError(msg)
    if rc == ST_ERR_INVALID_ARG:
        raise ValueError(msg)
    if rc == ST_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ST_ERR_LENGTH_MISMATCH:
        raise Exception(msg)
    raise SuchTreeError("CUDA path failed (no CPU fallback): " + msg)
