"""N>1 host logic on CPU: world_size-2 gloo run of the moment all-reduce and the
shard arithmetic (the GPU ranks run the same code with the nccl backend)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import REPO

from suchtree_b200 import _lib, shard


def _moments_of(x, y, x0, y0):
    m = _lib.Moments()
    m.n, m.x0, m.y0 = len(x), x0, y0
    m.sx, m.sy = float((x - x0).sum()), float((y - y0).sum())
    m.sxx, m.syy = float(((x - x0) ** 2).sum()), float(((y - y0) ** 2).sum())
    m.sxy = float(((x - x0) * (y - y0)).sum())
    return m


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n = 100001
    x = rng.random(n) * 40
    y = 0.5 * x + rng.random(n) * 10
    b, e = shard.pair_range(rank, world, n)
    m = _moments_of(x[b:e], y[b:e], 20.0, 15.0)
    shard.allreduce_moments(m)
    r = float(_lib.lib().st_moments_pearson(m))  # pure host arithmetic in the library
    q.put((rank, b, e, m.n, r))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_moment_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    x = rng.random(100001) * 40
    y = 0.5 * x + rng.random(100001) * 10
    want = float(np.corrcoef(x, y)[0, 1])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 100001
    for _, _, _, n, r in res:
        assert n == 100001 and r == pytest.approx(want, abs=1e-12)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 7, 64, 1000, 100000, 10**10])
def test_shards_partition_the_work(world, n):
    for fn, kw in ((shard.pair_range, {}), (shard.row_block, {})):
        edges = [fn(r, world, n, **kw) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
            assert e0 == b1 and b0 <= e0
        assert all(b % 2 == 0 or b == n for b, _ in edges) or fn is shard.row_block


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_balanced_shares(world):
    rng = np.random.default_rng(world)
    for n in (0, 1, 5, 1000):
        w = rng.integers(0, 50, n).astype(np.float64) ** 2
        shares = shard.balanced_shares(w, world)
        assert len(shares) == world
        assert np.array_equal(np.sort(np.concatenate(shares)) if n else np.concatenate(shares), np.arange(n))
        assert all(np.array_equal(s, np.sort(s)) for s in shares)
        loads = [w[s].sum() for s in shares]
        assert max(loads) - min(loads) <= (w.max() if n else 0)
        again = shard.balanced_shares(w, world)
        assert all(np.array_equal(a, b) for a, b in zip(shares, again))
