"""Multi-GPU decomposition of the path (one process per GPU, torch.distributed).

The index is replicated; work is partitioned with no data-path collective:
  * distances : contiguous ranges of the pair stream        -> pair_range()
  * matrix    : contiguous row blocks                        -> row_block()
  * sampler   : disjoint Philox sample ranges                -> pair_range()
  * clade scan: clades dealt out by link-pair count           -> balanced_shares()
The only exchange is the sampler's moment all-reduce (5 sums + n, fp64): NCCL over
NVLink on GPUs, gloo in the CPU tests.
"""
import ctypes as C


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
