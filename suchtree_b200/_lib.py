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
        ("paired_records", C.c_int32),
        ("probe_third_gather", C.c_double),
        ("probe_neighbour_hit", C.c_double),
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
    "st_build_id": (C.c_char_p, []),
    "st_pairs_kernel_id": (C.c_char_p, []),
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
    "st_quartet_topologies_device32": (_int, [_vp, _vp, _i64, _vp, _vp]),
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
    "st_host_route_info": (_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "st_host_pack_pairs": (_int, [_vp, _i64, _i64, _i64, _int, _vp, C.POINTER(C.c_uint64)]),
    "st_bench_copy": (_int, [_int, _i64, _i64, _i64, _int, C.POINTER(C.c_double)]),
    "st_host_alloc": (_int, [_i64, C.POINTER(_vp)]),
    "st_host_free": (_int, [_vp]),
    "st_host_trim": (_int, [_i64]),
    "st_host_register": (_int, [_vp, _i64]),
    "st_host_unregister": (_int, [_vp]),
    "st_host_is_pinned": (_int, [_vp]),
    "st_links_create": (_int, [_vp, _vp, _vp, _i64, C.POINTER(_vp)]),
    "st_links_destroy": (None, [_vp]),
    "st_links_linked_distances": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "st_links_sample_cycle": (_int, [_vp, C.POINTER(_u64), _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "st_links_sample_moments": (
        _int, [_vp, _u64, _i64, _i64, C.c_double, C.c_double, _vp, C.POINTER(Moments)]),
    "st_links_linked_moments": (
        _int, [_vp, _i64, _i64, C.c_double, C.c_double, _vp, C.POINTER(Moments)]),
    "st_links_clade_moments": (_int, [_vp, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "st_nccl_version": (_int, [C.POINTER(_int)]),
    "st_nccl_unique_id": (_int, [_vp]),
    "st_nccl_comm_create": (_int, [_int, _int, _int, _vp, C.POINTER(_vp)]),
    "st_nccl_comm_destroy": (_int, [_vp]),
}

# the bench-only library (include/suchtree_b200_bench.h): measurement tooling, not product
BENCH_SIGNATURES = {
    "st_bench_gather": (_int, [_int, _i64, _i64, _int, C.POINTER(C.c_double)]),
    "st_bench_last_error": (C.c_char_p, []),
}
_bench_lib = None


class _build_lock:
    """Inter-process lock (flock) around the stale-library check and rebuild."""

    def __enter__(self):
        import fcntl

        try:
            self.f = open(os.path.join(_build.HERE, ".build.lock"), "w")
            fcntl.flock(self.f, fcntl.LOCK_EX)
        except OSError:  # read-only install: nothing to build into anyway
            self.f = None
        return self

    def __exit__(self, *exc):
        if self.f is not None:
            import fcntl

            fcntl.flock(self.f, fcntl.LOCK_UN)
            self.f.close()
        return False


def _bind(L, signatures):
    for name, (res, args) in signatures.items():
        fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


def lib():
    """Load (building first if needed) the CUDA library.  Raises if impossible.
    A library whose embedded source hash (st_build_id) differs from the sources on
    disk is stale: it is rebuilt before loading, never silently used."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.LIB
                with _build_lock():  # one process per GPU: only one of them (re)builds, the rest wait
                    if _build.needs_build("product"):
                        _build.build(which=("product",))
                L = _bind(C.CDLL(path), SIGNATURES)
                if _build.sources("product") and L.st_build_id().decode() != _build.source_id("product"):
                    raise SuchTreeError(
                        "libsuchtree_b200.so (build id %s) does not match the sources on disk (%s)"
                        % (L.st_build_id().decode(), _build.source_id("product")))
                _lib = L
    return _lib


def bench_lib():
    """The bench-only library (gather roofline probe and experiment kernels)."""
    global _bench_lib
    if _bench_lib is None:
        with _lock:
            if _bench_lib is None:
                with _build_lock():
                    if _build.needs_build("bench"):
                        _build.build(which=("bench",))
                _bench_lib = _bind(C.CDLL(_build.BENCH_LIB), BENCH_SIGNATURES)
    return _bench_lib


_py_glue = False  # False: not tried yet; None: unavailable


def py_glue():
    """The CPython glue library (name -> id walks in C), or None when it cannot be built here
    (no compiler / no Python.h): callers keep their Python walk then.  Loaded with PyDLL: the
    GIL stays held, arguments are Python objects."""
    global _py_glue
    if _py_glue is False:
        with _lock:
            if _py_glue is False:
                g = None
                try:
                    with _build_lock():
                        if _build.py_glue_needs_build():
                            _build.build_py_glue()
                    if os.path.exists(_build.PY_LIB):
                        g = C.PyDLL(_build.PY_LIB)
                        g.st_py_names_to_ids.restype = C.c_ssize_t
                        g.st_py_names_to_ids.argtypes = [C.py_object, C.py_object, C.c_void_p, C.c_ssize_t]
                except Exception:
                    g = None
                _py_glue = g
    return _py_glue


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


# --------------------------------------------------------------------------- #
# page-locked result arrays and registered inputs
# --------------------------------------------------------------------------- #
class _PinnedBlock:
    """A block of the library's pinned pool (st_host_alloc) exposed through the array
    interface; it goes back to the pool when the last array viewing it dies."""

    __slots__ = ("ptr", "__array_interface__", "__weakref__")

    def __init__(self, shape, dtype):
        import numpy as np

        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        p = _vp()
        self.ptr = None
        check(lib().st_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.__array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": dt.str,
                                    "data": (self.ptr, False), "version": 3}

    def __del__(self):
        if self.ptr:
            try:
                lib().st_host_free(self.ptr)
            except Exception:  # interpreter shutdown
                pass
            self.ptr = None


def result_empty(shape, dtype, zero_small=False):
    """The fresh result array of a drop-in call: page-locked (pool) when it is large, an ordinary
    np.empty / np.zeros otherwise.  SUCHTREE_B200_PINNED_RESULTS=0 keeps every result pageable
    (slower D2H, but no page-locked memory is held by arrays the caller keeps around)."""
    import numpy as np

    if not isinstance(shape, tuple):
        shape = (int(shape),)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    if nbytes >= PINNED_RESULT_MIN_BYTES and os.environ.get("SUCHTREE_B200_PINNED_RESULTS", "1") != "0":
        return pinned_empty(shape, dtype)
    return np.zeros(shape, dtype=dtype) if zero_small else np.empty(shape, dtype=dtype)


def pinned_empty(shape, dtype):
    """np.empty() in page-locked memory from the library's pool: an ordinary ndarray
    (its .base keeps the block alive) that D2H copies can land in directly."""
    import numpy as np

    if not isinstance(shape, tuple):
        shape = (int(shape),)
    return np.asarray(_PinnedBlock(shape, dtype))


PINNED_RESULT_MIN_BYTES = 4 << 20    # smaller results: plain np.empty (the small-call path copies anyway)
REGISTER_MIN_BYTES = 64 << 20        # inputs at least this large are candidates for registration

_registered = {}   # id(owner) -> [ptr, nbytes, sightings, registered, finalizer]
# re-entrant: a finaliser (_unregister) can fire from a garbage collection that an allocation
# inside maybe_register() triggers while this thread already holds the lock
_reg_lock = threading.RLock()


def _register_policy():
    """SUCHTREE_B200_REGISTER: 0 = never page-lock caller arrays, 1 = the first time a large
    input is seen, 2 (default) = the second time the same array is seen (a one-off call
    should not pay ~0.2 s/GB of pinning; a loop over the same array should)."""
    try:
        return int(os.environ.get("SUCHTREE_B200_REGISTER", "2"))
    except ValueError:
        return 2


def _unregister(key, ptr):
    with _reg_lock:
        ent = _registered.pop(key, None)
    if ent is not None and ent[3]:
        try:
            lib().st_host_unregister(ptr)
        except Exception:
            pass


def maybe_register(arr):
    """Called with a large C-contiguous input array: page-lock the buffer of the ndarray
    that owns the memory once the policy says so.  The registration lives exactly as long
    as the owner (weakref finaliser), so a freed-and-reused address is never mistaken
    for a registered one.  Returns True when the array is (now) page-locked."""
    import weakref

    import numpy as np

    policy = _register_policy()
    if policy <= 0 or arr.nbytes < REGISTER_MIN_BYTES:
        return False
    owner = arr
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None or not owner.flags.owndata:
        return False  # memory owned by something we cannot track (mmap, bytes, another library)
    ptr, nbytes, key = owner.ctypes.data, owner.nbytes, id(owner)
    with _reg_lock:
        ent = _registered.get(key)
        if ent is not None and (ent[0] != ptr or ent[1] != nbytes):
            ent = None  # the owner was resized in place: forget the old range
            stale = _registered.pop(key)
            if stale[3]:
                lib().st_host_unregister(stale[0])
            stale[4].detach()
        if ent is None:
            try:
                fin = weakref.finalize(owner, _unregister, key, ptr)
            except TypeError:
                return False
            ent = _registered[key] = [ptr, nbytes, 0, False, fin]
        ent[2] += 1
        if ent[3]:
            return True
        if ent[2] >= policy:
            if lib().st_host_register(ptr, nbytes) == ST_OK:
                ent[3] = True
                return True
            ent[2] = -(1 << 60)  # registration failed (locked-memory limit ...): do not retry every call
    return False
