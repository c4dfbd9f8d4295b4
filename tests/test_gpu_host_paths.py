"""The host-buffer entry points (what the drop-in Python call uses) across their
internal routes: the small-call zero-copy path, the packed (bit-stream staging) path, the
hybrid packed/direct split for pinned input, pageable vs pinned results, strided input,
chunk boundaries, and the range errors each route must report like the reference
(MuchTree.pyx:897-903)."""
import os

import numpy as np
import pytest

import oracle as O
from suchtree_b200 import InvalidNodeError, SuchTree, synth

pytestmark = pytest.mark.gpu

CHUNK = 1 << 22  # ST_STAGE_PAIRS_MAX


@pytest.fixture(scope="module")
def tree():
    ft = synth.yule_tree(30000, seed=12)
    return SuchTree.from_flat(ft), O.OracleTree(ft.parent, ft.distance), ft


def _pairs(ft, n, seed):
    return np.random.default_rng(seed).integers(0, ft.size, size=(n, 2)).astype(np.int64)


@pytest.mark.parametrize("n", [1, 2, 3, 4095, 4096, 4097, 100001])
def test_small_and_medium_calls_match_oracle(tree, n):
    T, ot, ft = tree
    p = _pairs(ft, n, n)
    assert np.array_equal(T.distances_bulk(p), ot.distances_f64_climb(p))
    assert np.array_equal(T.common_ancestors_bulk(p), ot.distances_f64_climb(p, with_mrca=True)[1])


def test_multi_chunk_pageable_pinned_and_hybrid_agree(tree):
    """2 chunks + a ragged tail through every combination of pageable / pinned input and
    output; pinned contiguous input takes the hybrid split (part bit-packed, part int64)."""
    import torch

    T, ot, ft = tree
    n = 2 * CHUNK + 12345
    p = _pairs(ft, n, 99)
    want = T.distances_bulk(p)  # pageable in, pageable out
    sel = np.random.default_rng(1).integers(0, n, size=200000)
    assert np.array_equal(want[sel], ot.distances_f64_climb(p[sel]))
    hp = torch.from_numpy(p).pin_memory()
    ho = torch.empty(n, dtype=torch.float64).pin_memory()
    got = T.distances_bulk(hp.numpy(), out=ho.numpy())  # pinned in (hybrid), pinned out
    assert got is not None and np.array_equal(ho.numpy(), want)
    assert np.array_equal(T.distances_bulk(hp.numpy()), want)  # pinned in, pageable out
    assert np.array_equal(T.distances_bulk(p, out=ho.numpy().copy()), want)  # pageable in, given out
    for frac in ("0", "1", "0.3"):
        os.environ["SUCHTREE_B200_PACK_FRACTION"] = frac
        try:
            assert np.array_equal(T.distances_bulk(hp.numpy()), want), frac
        finally:
            del os.environ["SUCHTREE_B200_PACK_FRACTION"]
    os.environ["SUCHTREE_B200_HOST_PATH"] = "direct"
    try:
        assert np.array_equal(T.distances_bulk(hp.numpy()), want)
    finally:
        del os.environ["SUCHTREE_B200_HOST_PATH"]
    m = T.common_ancestors_bulk(hp.numpy())
    assert np.array_equal(m[sel], ot.distances_f64_climb(p[sel], with_mrca=True)[1])


def test_strided_inputs_on_the_packed_path(tree):
    T, ot, ft = tree
    n = 300000
    base = np.zeros((n, 6), dtype=np.int64)
    p = _pairs(ft, n, 5)
    base[:, 1] = p[:, 0]
    base[:, 4] = p[:, 1]
    view = base[:, 1:5:3]
    assert view.strides == (48, 24)
    want = ot.distances_f64_climb(p)
    assert np.array_equal(T.distances_bulk(view), want)
    rev = view[::-1]
    assert np.array_equal(T.distances_bulk(rev), want[::-1])
    tr = np.ascontiguousarray(p.T).T  # Fortran-ordered (n,2)
    assert np.array_equal(T.distances_bulk(tr), want)


@pytest.mark.parametrize("n", [7, 5000, CHUNK + 10])
def test_range_errors_on_every_route(tree, n):
    """max id if it is >= size, else min id -- including ids that do not fit int32,
    which the packed path must catch before narrowing."""
    import torch

    T, _, ft = tree
    for pinned in (False, True):
        def arr(p):
            return torch.from_numpy(p).pin_memory().numpy() if pinned else p

        p = _pairs(ft, n, 3)
        p[n // 2, 1] = ft.size + 5
        with pytest.raises(InvalidNodeError) as e:
            T.distances_bulk(arr(p))
        assert e.value.node_id == ft.size + 5 and e.value.tree_size == ft.size
        p = _pairs(ft, n, 4)
        p[n - 1, 0] = (1 << 40) + 3  # low 32 bits alone would be a valid id
        with pytest.raises(InvalidNodeError) as e:
            T.distances_bulk(arr(p))
        assert e.value.node_id == (1 << 40) + 3
        p = _pairs(ft, n, 5)
        p[0, 0] = -(1 << 33)
        p[n // 3, 1] = -2
        with pytest.raises(InvalidNodeError) as e:
            T.common_ancestors_bulk(arr(p))
        assert e.value.node_id == -(1 << 33)
        p[n // 2, 0] = ft.size  # too large and negative present: the max wins
        with pytest.raises(InvalidNodeError) as e:
            T.distances_bulk(arr(p))
        assert e.value.node_id == ft.size
        # and the tree still answers
        q = _pairs(ft, 10, 6)
        assert np.all(np.isfinite(T.distances_bulk(q)))


def test_concurrent_host_calls_from_threads(tree):
    """Handles are immutable after creation: host threads may query one tree (and
    several trees) concurrently; the GIL is released inside the C-ABI calls."""
    import threading

    T, ot, ft = tree
    ft2 = synth.balanced_tree(4096, seed=3)
    T2, ot2 = SuchTree.from_flat(ft2), O.OracleTree(ft2.parent, ft2.distance)
    jobs = []
    for k in range(8):
        tr, orc, f = (T, ot, ft) if k % 2 == 0 else (T2, ot2, ft2)
        n = [3, 5000, 300000, 70000][k % 4]
        p = np.random.default_rng(100 + k).integers(0, f.size, size=(n, 2)).astype(np.int64)
        jobs.append((tr, orc, p))
    results = [None] * len(jobs)

    def run(i):
        tr, _, p = jobs[i]
        for _ in range(3):
            results[i] = (tr.distances_bulk(p), tr.common_ancestors_bulk(p))

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
        assert not th.is_alive()
    for (tr, orc, p), (d, m) in zip(jobs, results):
        want_d, want_m = orc.distances_f64_climb(p, with_mrca=True)
        assert np.array_equal(d, want_d) and np.array_equal(m, want_m)


@pytest.mark.parametrize("kind", ["wide", "tables", "compact_small_blocks", "caterpillar", "three_nodes"])
def test_bit_packed_stream_on_every_layout_and_id_width(kind):
    """The chunked host pipeline ships ids as a 2 x ceil(log2 n_nodes)-bit stream that the pair
    kernel decodes; every index layout has its own instantiation of that kernel, and the width
    runs from 2 bits (a 3-node tree) up.  Distances and MRCA ids against the oracle, ragged
    counts (whole groups of 64 pairs, a partial last word, the odd last pair)."""
    if kind == "three_nodes":
        ft = synth.yule_tree(2, seed=1)
    elif kind == "caterpillar":
        ft = synth.caterpillar_tree(6000, seed=2)
    else:
        ft = synth.yule_tree(20000, seed=13)
    ot = O.OracleTree(ft.parent, ft.distance)
    env = {"tables": {"SUCHTREE_B200_LAYOUT": "tables"}}.get(kind, {})
    os.environ.update(env)
    try:
        T = SuchTree.from_flat(ft, _wide=True) if kind == "wide" else (
            SuchTree.from_flat(ft, _block_shift=3) if kind == "compact_small_blocks" else SuchTree.from_flat(ft))
    finally:
        for k in env:
            del os.environ[k]
    for n in (300_001, 262_145 + 63, 1 << 19):
        p = _pairs(ft, n, n % 1000)
        want, wm = ot.distances_f64_climb(p, with_mrca=True)
        assert np.array_equal(T.distances_bulk(p), want), (kind, n)
        assert np.array_equal(T.common_ancestors_bulk(p), wm), (kind, n)
