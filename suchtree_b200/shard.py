"""Multi-GPU decomposition of the path (one process per GPU, torch.distributed).

The index is replicated; work is partitioned with no data-path collective:
  * distances : contiguous ranges of the pair stream        -> pair_range()
  * matrix    : contiguous row blocks                        -> row_block()
  * sampler   : disjoint Philox sample ranges                -> pair_range()
  * clade scan: clades dealt out by link-pair count           -> balanced_shares()
The only exchange is the sampler's moment all-reduce (5 sums + n, fp64): NCCL over
NVLink on GPUs -- enqueued by the library on the moment kernel's own stream
(MomentComm -> st_links_sample_moments(..., nccl_comm)) -- gloo in the CPU tests
(allreduce_moments).
"""
import ctypes as C


class MomentComm:
    """The library's own NCCL communicator for the moment all-reduce (one per process,
    one process per GPU).  The 128-byte ncclUniqueId is made on rank 0 and handed to the
    other ranks through torch.distributed (any backend: it is plumbing only)."""

    def __init__(self, device, rank=None, world=None, group=None):
        import numpy as np
        import torch
        import torch.distributed as dist

        from . import _lib

        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.world = dist.get_world_size(group) if world is None else int(world)
        ident = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _lib.check(_lib.lib().st_nccl_unique_id(ident.ctypes.data))
        if dist.get_backend(group) == "nccl":
            t = torch.from_numpy(ident).to(torch.device("cuda", int(device)))
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu().numpy()
        else:
            t = torch.from_numpy(ident)
            dist.broadcast(t, src=0, group=group)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().st_nccl_comm_create(int(device), self.world, self.rank, ident.ctypes.data,
                                                  C.byref(self.handle)))

    def close(self):
        from . import _lib

        if self.handle is not None and self.handle.value:
            _lib.lib().st_nccl_comm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pair_range(rank, world, n_total, align=2):
    """[begin, end) of rank's contiguous share of n_total items; begin is a multiple of
    `align` (the Philox generators consume the stream two items per call)."""
    per = -(-n_total // world)
    per += (-per) % align
    b = min(rank * per, n_total)
    return b, min(b + per, n_total)


def row_block(rank, world, n_rows, tile=64):
    """[begin, end) rows of rank's block of an n_rows matrix, tile-aligned."""
    per = -(-n_rows // world)
    per += (-per) % tile
    b = min(rank * per, n_rows)
    return b, min(b + per, n_rows)


def balanced_shares(weights, world):
    """Indices of the items each rank takes so that the summed weights come out close:
    items sorted by weight (descending, ties by index) and dealt out boustrophedon
    (0..world-1, world-1..0, ...).  Deterministic; every index appears exactly once; each
    share is returned in ascending index order."""
    import numpy as np

    w = np.asarray(weights, dtype=np.float64)
    order = np.lexsort((np.arange(w.shape[0]), -w))
    pos = np.arange(w.shape[0])
    lap, k = np.divmod(pos, world)
    owner = np.where(lap % 2 == 0, k, world - 1 - k)
    return [np.sort(order[owner == r]) for r in range(world)]


def allreduce_moments(m, group=None, device=None):
    """Sum the shifted moments of all ranks in place (x0, y0 must be the same on every
    rank) and return the struct.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return m
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([m.n, m.sx, m.sy, m.sxx, m.syy, m.sxy], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    m.n, m.sx, m.sy, m.sxx, m.syy, m.sxy = (float(v) for v in t.tolist())
    return m
