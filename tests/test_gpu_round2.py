"""Round-2 behaviour of the drop-in boundary on a real GPU:

* distances_bulk(pageable int64 array) -> FRESH float64 array, the reference's signature
  (MuchTree.pyx:872-909), served from the page-locked result pool, with large repeated
  inputs registered in place -- same numbers as the oracle whichever route is taken;
* every host call owns its range-status word: concurrent callers on one tree neither
  steal nor inherit each other's InvalidNodeError (MuchTree.pyx:897-903), and a flag left
  behind by the device API never leaks into a host call;
* the link-list handle (st_links), the fixed-order bucket statistics of
  sample_linked_distances (MuchTree.pyx:3045-3067), the NCCL moment all-reduce;
* the matrix writer at BASELINE cfg5's full size against the oracle.
"""
import ctypes as C
import gc
import json
import os
import threading

import numpy as np
import pytest

import oracle as O
from conftest import DATA, GOLDEN
from suchtree_b200 import InvalidNodeError, SuchLinkedTrees, SuchTree, _lib, shard, synth
from suchtree_b200.linked import moments_pearson

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tree():
    ft = synth.yule_tree(40000, seed=21)
    return SuchTree.from_flat(ft), O.OracleTree(ft.parent, ft.distance), ft


def _pairs(ft, n, seed):
    return np.random.default_rng(seed).integers(0, ft.size, size=(n, 2)).astype(np.int64)


def _pinned(a):
    return bool(_lib.lib().st_host_is_pinned(a.ctypes.data))


# ------------------------------------------------------------ result pool ----
def test_fresh_result_comes_from_the_pinned_pool_and_is_recycled(tree):
    T, ot, ft = tree
    n = 3_000_000
    p = _pairs(ft, n, 1)
    gc.collect()
    _lib.lib().st_host_trim(0)  # blocks cached by earlier tests would be handed out first
    r = T.distances_bulk(p)
    assert type(r) is np.ndarray and r.dtype == np.float64 and r.shape == (n,) and r.flags.c_contiguous
    assert r.flags.writeable and _pinned(r)
    sel = np.random.default_rng(2).integers(0, n, 100000)
    assert np.array_equal(r[sel], ot.distances_f64_climb(p[sel]))
    keep = r.copy()
    ptr = r.ctypes.data
    view = r[10:20]  # a view keeps the block alive
    del r
    gc.collect()
    r2 = T.distances_bulk(p)
    assert r2.ctypes.data != ptr and np.array_equal(view, keep[10:20])
    del view
    gc.collect()
    r3 = T.distances_bulk(p)  # the first block is back in the pool: same size class -> recycled
    assert r3.ctypes.data == ptr
    assert np.array_equal(r2, keep) and np.array_equal(r3, keep)
    # small results are ordinary arrays
    s = T.distances_bulk(p[:1000])
    assert s.flags.owndata and np.array_equal(s, keep[:1000])


def test_pool_trim_and_foreign_pointer():
    L = _lib.lib()
    a = _lib.pinned_empty((1 << 20,), np.float64)
    assert _pinned(a)
    a[:] = 1.5
    del a
    gc.collect()
    assert L.st_host_trim(0) == 0
    x = np.zeros(4)
    assert L.st_host_free(x.ctypes.data) == _lib.ST_ERR_INVALID_ARG


def test_large_pool_blocks_are_huge_page_mappings_locked_in_place():
    """Blocks of >= 4 MiB are anonymous mappings (2 MiB aligned, transparent huge pages where the
    kernel grants them) page-locked with cudaHostRegister; smaller ones come from cudaHostAlloc.
    Both kinds are page-locked, writable everywhere, recycled by size class and released by trim."""
    L = _lib.lib()
    gc.collect()
    L.st_host_trim(0)
    big = _lib.pinned_empty((5 << 20,), np.float64)  # 40 MiB
    small = _lib.pinned_empty((1 << 17,), np.float64)  # 1 MiB
    assert _pinned(big) and _pinned(small) and _pinned(big[-1:])
    if os.environ.get("SUCHTREE_B200_PINNED_MMAP", "1") != "0":
        assert big.ctypes.data % (2 << 20) == 0
    big[:] = 2.5
    big[-1] = 7.0
    small[:] = 1.0
    assert big[0] == 2.5 and big[-1] == 7.0 and float(small.sum()) == float(1 << 17)
    ptr = big.ctypes.data
    del big
    gc.collect()
    again = _lib.pinned_empty((5 << 20,), np.float64)
    assert again.ctypes.data == ptr  # same size class: the cached block
    del again, small
    gc.collect()
    assert L.st_host_trim(0) == 0
    fresh = _lib.pinned_empty((5 << 20,), np.float64)  # mapped and locked anew after the trim
    assert _pinned(fresh)
    fresh[::4096] = 1.0


# ------------------------------------------------------------ registration ---
def test_large_repeated_input_is_registered_in_place_then_released(tree):
    T, ot, ft = tree
    n = (_lib.REGISTER_MIN_BYTES // 16) + 1001  # just over the threshold
    p = _pairs(ft, n, 3)
    assert not _pinned(p)
    r1 = T.distances_bulk(p)
    assert not _pinned(p)  # default policy: a one-off call does not pay for pinning
    r2 = T.distances_bulk(p)
    assert _pinned(p)  # second sighting: page-locked in place, DMA'd directly from now on
    r3 = T.distances_bulk(p)
    r4 = T.distances_bulk(p[: n // 2])  # a view of a registered array
    sel = np.random.default_rng(4).integers(0, n, 100000)
    want = ot.distances_f64_climb(p[sel])
    for r in (r1, r2, r3):
        assert np.array_equal(r[sel], want)
    assert np.array_equal(r4, r1[: n // 2])
    key = id(p)
    assert key in _lib._registered
    del p
    gc.collect()
    assert key not in _lib._registered  # unregistered with the array


def test_register_policy_env(tree):
    T, ot, ft = tree
    n = (_lib.REGISTER_MIN_BYTES // 16) + 7
    p = _pairs(ft, n, 5)
    os.environ["SUCHTREE_B200_REGISTER"] = "0"
    try:
        a = T.distances_bulk(p)
        b = T.distances_bulk(p)
        assert not _pinned(p)
    finally:
        del os.environ["SUCHTREE_B200_REGISTER"]
    os.environ["SUCHTREE_B200_REGISTER"] = "1"
    try:
        c = T.distances_bulk(p)
        assert _pinned(p)
    finally:
        del os.environ["SUCHTREE_B200_REGISTER"]
    assert np.array_equal(a, b) and np.array_equal(a, c)


# ------------------------------------------- per-call status word, lanes -----
def test_concurrent_callers_keep_their_own_range_errors(tree):
    """Four threads on ONE tree: two with valid pairs, two with an out-of-range id each (one
    too large, one negative), medium and large calls mixed.  Each bad caller gets ITS error
    (the id the reference would report), the good ones get correct distances."""
    T, ot, ft = tree
    n_big, n_med = 5_000_000, 100_000
    good_big, good_med = _pairs(ft, n_big, 6), _pairs(ft, n_med, 7)
    bad_hi = _pairs(ft, n_big, 8)
    bad_hi[n_big - 5, 1] = ft.size + 123
    bad_lo = _pairs(ft, n_med, 9)
    bad_lo[17, 0] = -44
    results, errors = {}, {}

    def run(name, arr):
        for rep in range(3):
            try:
                results[(name, rep)] = T.distances_bulk(arr)
            except Exception as e:  # noqa: BLE001
                errors[(name, rep)] = e

    th = [threading.Thread(target=run, args=a) for a in
          (("good_big", good_big), ("bad_hi", bad_hi), ("good_med", good_med), ("bad_lo", bad_lo))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for rep in range(3):
        e = errors.get(("bad_hi", rep))
        assert isinstance(e, InvalidNodeError) and e.node_id == ft.size + 123, e
        e = errors.get(("bad_lo", rep))
        assert isinstance(e, InvalidNodeError) and e.node_id == -44, e
        assert ("good_big", rep) not in errors and ("good_med", rep) not in errors, errors
    sel = np.random.default_rng(10).integers(0, n_big, 50000)
    assert np.array_equal(results[("good_big", 2)][sel], ot.distances_f64_climb(good_big[sel]))
    assert np.array_equal(results[("good_med", 2)], ot.distances_f64_climb(good_med))


def test_device_api_flag_does_not_leak_into_host_calls(tree):
    import torch

    T, ot, ft = tree
    bad = torch.tensor([[0, ft.size + 5], [2, 4]], dtype=torch.int32, device="cuda")
    out = torch.empty(2, dtype=torch.float64, device="cuda")
    T.distances_device(bad.data_ptr(), 2, out.data_ptr(), idx_bits=32)
    torch.cuda.synchronize()
    # a stale device-API flag is pending; valid host calls of every size class must not see it
    for n in (10, 100_000, 5_000_000):
        p = _pairs(ft, n, n)
        r = T.distances_bulk(p)
        assert np.array_equal(r[:1000], ot.distances_f64_climb(p[:1000]))
    q = 2 * np.random.default_rng(0).integers(0, ft.n_leaves, size=(300_000, 4))
    assert T.quartet_topologies_bulk(q.astype(np.int64)).shape == (300_000, 4)
    D = T.pairwise_distances(list(range(0, 200, 2)))
    assert D.shape == (100, 100)
    # ... and the device API still reports its own
    with pytest.raises(InvalidNodeError) as ei:
        T.check_range()
    assert ei.value.node_id == ft.size + 5
    T.check_range()  # cleared


def test_matrix_host_id_errors_are_private_too(tree):
    T, ot, ft = tree
    with pytest.raises(InvalidNodeError):
        T.pairwise_distances([0, 2, ft.size + 9])
    D = T.pairwise_distances([0, 2, 4])
    assert D[0, 0] == 0.0 and D[0, 1] == T.distance(0, 2)


# ------------------------------------------------------------------ matrix ---
def test_matrix_writer_covers_every_element(tree):
    """The host path hands the writer recycled (dirty) pool memory: every element of a row
    block must be written -- checked on a NaN-filled device buffer, ragged sizes."""
    import torch

    T, ot, ft = tree
    for n, r0, r1 in ((ft.n_leaves, 0, 700), (ft.n_leaves, 39000, 40000), (ft.n_leaves, 12345, 12500)):
        buf = torch.full((r1 - r0, n), float("nan"), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().st_distance_matrix(T._handle, None, n, r0, r1, buf.data_ptr(), 1, None))
        torch.cuda.synchronize()
        assert not bool(torch.isnan(buf).any().item())
        rows = np.array([0, (r1 - r0) // 2, r1 - r0 - 1])
        for rr in rows:
            p = np.stack([np.full(n, 2 * (r0 + rr)), 2 * np.arange(n)], axis=1).astype(np.int64)
            assert np.array_equal(buf[rr].cpu().numpy(), ot.distances_f64_climb(p))


def test_pairwise_distances_large_host_result_is_pinned_and_exact(tree):
    T, ot, ft = tree
    nodes = list(range(0, 6000, 2))  # 3000 x 3000 x 8 B = 72 MB: several bands
    D = T.pairwise_distances(nodes)
    assert D.shape == (3000, 3000) and _pinned(D)
    assert np.array_equal(D, D.T) and np.all(np.diagonal(D) == 0.0)
    a, b = np.meshgrid(nodes[:50], nodes, indexing="ij")
    want = ot.distances_f64_climb(np.stack([a.ravel(), b.ravel()], axis=1).astype(np.int64)).reshape(50, 3000)
    assert np.array_equal(D[:50], want)
    # unsorted id list with internal nodes and a repeat: the generic kernel, same host path
    rng = np.random.default_rng(3)
    ids = rng.integers(0, ft.size, 900).tolist() + [5, 5]
    D2 = T.pairwise_distances(ids)
    a, b = np.meshgrid(ids, ids, indexing="ij")
    want = ot.distances_f64_climb(np.stack([a.ravel(), b.ravel()], axis=1).astype(np.int64)).reshape(902, 902)
    assert np.array_equal(D2, want)


@pytest.mark.timeout(900)
def test_cfg5_full_size_matrix_against_oracle():
    """BASELINE cfg5 at full size: the 100,000-leaf Yule tree, all 8 row shards of the
    100k x 100k matrix (10 GB each, one after another through one buffer): 1e6 sampled
    elements -- uniform, diagonal tiles, both sides of the diagonal, last partial tile --
    and whole rows with their checksums, bit for bit against O2."""
    import torch

    import bench

    ft = synth.yule_tree(bench.TREE_LEAVES, seed=bench.TREE_SEED)
    T = SuchTree.from_flat(ft)
    n = bench.TREE_LEAVES
    rb, re_ = shard.row_block(0, 8, n)
    block = torch.empty((re_ - rb, n), dtype=torch.float64, device="cuda")

    def mat(r0, r1):
        _lib.check(_lib.lib().st_distance_matrix(T._handle, None, n, r0, r1, block.data_ptr(), 1, None))

    res = bench.cfg5_parity_vs_oracle(ft, T, block, n, 8, mat, torch.device("cuda"))
    assert res["sampled_elements"] >= 1_000_000 - 8 and res["whole_rows"] >= 16
    assert res["bit_exact"], res


# ------------------------------------------------------------ linked trees ---
@pytest.fixture(scope="module")
def linked():
    fa, fb = synth.yule_tree(3000, seed=31, names=True), synth.yule_tree(4000, seed=32, names=True)
    TA, TB = SuchTree.from_flat(fa), SuchTree.from_flat(fb)
    rng = np.random.default_rng(33)
    ll = np.stack([2 * rng.integers(0, 4000, 2500), 2 * rng.integers(0, 3000, 2500)], axis=1).astype(np.int64)
    return TA, TB, np.ascontiguousarray(ll)


def test_links_handle_matches_raw_linklist_calls(linked):
    TA, TB, ll = linked
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.st_links_create(TA._handle, TB._handle, ll.ctypes.data, ll.shape[0], C.byref(h)))
    try:
        for first, cnt in ((0, 100001), (12346, 77777)):
            m1, m2 = _lib.Moments(), _lib.Moments()
            _lib.check(L.st_sample_moments(TA._handle, TB._handle, ll.ctypes.data, ll.shape[0], 7, first, cnt,
                                           0.5, 0.25, C.byref(m1)))
            _lib.check(L.st_links_sample_moments(h, 7, first, cnt, 0.5, 0.25, None, C.byref(m2)))
            assert bytes(m1) == bytes(m2)
            _lib.check(L.st_linked_moments(TA._handle, TB._handle, ll.ctypes.data, ll.shape[0], first, cnt,
                                           0.5, 0.25, C.byref(m1)))
            _lib.check(L.st_links_linked_moments(h, first, cnt, 0.5, 0.25, None, C.byref(m2)))
            assert bytes(m1) == bytes(m2)
        # shards of the sample range add up to the whole (what the all-reduce relies on)
        whole = _lib.Moments()
        _lib.check(L.st_links_sample_moments(h, 9, 0, 400000, 1.0, 1.0, None, C.byref(whole)))
        parts = []
        for r in range(4):
            m = _lib.Moments()
            _lib.check(L.st_links_sample_moments(h, 9, r * 100000, 100000, 1.0, 1.0, None, C.byref(m)))
            parts.append(m)
        tot = _lib.Moments(sum(m.n for m in parts), 1.0, 1.0, *(sum(getattr(m, k) for m in parts)
                                                                 for k in ("sx", "sy", "sxx", "syy", "sxy")))
        assert tot.n == whole.n
        assert abs(moments_pearson(tot) - moments_pearson(whole)) < 1e-12
    finally:
        L.st_links_destroy(h)
    bad = ll.copy()
    bad[5, 1] = TA.size + 1
    assert L.st_links_create(TA._handle, TB._handle, bad.ctypes.data, bad.shape[0], C.byref(h)) == _lib.ST_ERR_NODE_RANGE


def test_linked_object_reuses_its_handle_until_the_subset_changes(linked):
    TA, TB, ll = linked
    S = SuchLinkedTrees.from_linklist(TA, TB, ll)
    r0 = S.sample_pearson(200000, seed=3)
    h0 = S._links().value
    assert S.sample_pearson(200000, seed=3) == r0 and S._links().value == h0
    want = moments_pearson(S.linked_moments())
    S.subset_b(TB.root_node)  # same links, new subset version -> a new handle
    assert abs(moments_pearson(S.linked_moments()) - want) < 1e-12
    ld = S.linked_distances()
    assert ld["n_pairs"] == 2500 * 2499 // 2
    pa = O.linked_pairs(S.linklist)
    ota = O.OracleTree(TA._ft.parent, TA._ft.distance)
    assert np.array_equal(ld["ids_A"], pa[0]) and np.array_equal(ld["TreeA"][:5000], ota.distances_f64(pa[0][:5000]))


def test_sampler_bucket_statistics_are_reproducible_bit_for_bit(linked):
    """The reference's bucket sums are sequential and reproducible; ours are reduced in a
    fixed order on the device (no floating-point atomics): two runs from the same numpy seed
    give identical deviations, sample counts and samples."""
    TA, TB, ll = linked
    runs = []
    for _ in range(3):
        np.random.seed(1234)
        S = SuchLinkedTrees.from_linklist(TA, TB, ll)
        runs.append(S.sample_linked_distances(sigma=0.05, buckets=16, n=512, maxcycles=50))
    assert runs[0] is not None
    for r in runs[1:]:
        assert r["n_samples"] == runs[0]["n_samples"]
        assert r["deviation_a"] == runs[0]["deviation_a"] and r["deviation_b"] == runs[0]["deviation_b"]
        assert np.array_equal(r["TreeA"], runs[0]["TreeA"]) and np.array_equal(r["TreeB"], runs[0]["TreeB"])
    # bucket sums equal a host restatement of MuchTree.pyx:3045-3052 to rounding
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.st_links_create(TA._handle, TB._handle, ll.ctypes.data, ll.shape[0], C.byref(h)))
    seed = C.c_uint64(99)
    oa, ob = np.empty(16 * 512), np.empty(16 * 512)
    s = [np.zeros(16) for _ in range(4)]
    _lib.check(L.st_links_sample_cycle(h, C.byref(seed), 16, 512, oa.ctypes.data, ob.ctypes.data,
                                       s[0].ctypes.data, s[1].ctypes.data, s[2].ctypes.data, s[3].ctypes.data))
    L.st_links_destroy(h)
    assert np.allclose(s[0], oa.reshape(16, 512).sum(1), rtol=1e-13)
    assert np.allclose(s[1], (oa ** 2).reshape(16, 512).sum(1), rtol=1e-13)
    assert np.allclose(s[2], ob.reshape(16, 512).sum(1), rtol=1e-13)
    assert np.allclose(s[3], (ob ** 2).reshape(16, 512).sum(1), rtol=1e-13)


def test_sample_cycle_rejects_runs_beyond_the_jump_table(linked):
    TA, TB, ll = linked
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.st_links_create(TA._handle, TB._handle, ll.ctypes.data, ll.shape[0], C.byref(h)))
    seed = C.c_uint64(1)
    x = np.zeros(4)
    rc = L.st_links_sample_cycle(h, C.byref(seed), 1 << 30, 1 << 30, x.ctypes.data, x.ctypes.data, x.ctypes.data,
                                 x.ctypes.data, x.ctypes.data, x.ctypes.data)
    L.st_links_destroy(h)
    assert rc == _lib.ST_ERR_INVALID_ARG


# ---------------------------------------------------------- thin callers -----
def test_get_descendants_and_minus_one_edges_match_the_reference(golden_trees):
    with open(os.path.join(GOLDEN, "r2.json")) as f:
        g = json.load(f)
    for name, want in g["descendants"].items():
        rec = golden_trees[name]
        T = SuchTree(rec.get("newick") or os.path.join(DATA, name))
        gen = T.get_descendants(T.root_node)
        assert next(gen) == T.root_node  # a generator, start node first
        for i, w in enumerate(want):
            assert list(T.get_descendants(i)) == w, (name, i)
    T = SuchTree(g["minus_one"]["newick"])
    got = [T.distance_to_root(i) for i in range(T.size)]
    assert got == g["minus_one"]["distance_to_root"]


# -------------------------------------------------------------- NCCL ---------
def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only: the id broadcast
    try:
        fa, fb = synth.yule_tree(3000, seed=31, names=True), synth.yule_tree(4000, seed=32, names=True)
        TA, TB = SuchTree.from_flat(fa, device=rank), SuchTree.from_flat(fb, device=rank)
        rng = np.random.default_rng(33)
        ll = np.stack([2 * rng.integers(0, 4000, 2500), 2 * rng.integers(0, 3000, 2500)], axis=1).astype(np.int64)
        S = SuchLinkedTrees.from_linklist(TA, TB, ll)
        comm = shard.MomentComm(rank)
        n = 1_000_000
        r_all = S.sample_pearson(n, seed=5, comm=comm)
        m = S.sample_moments(n // world, seed=5, first_sample=rank * (n // world), comm=comm)
        r_one = S.sample_pearson(n, seed=5) if rank == 0 else None
        total = 2500 * 2499 // 2
        b, e = shard.pair_range(rank, world, total)
        mx = S.linked_moments(b, e - b, comm=comm)
        r_ex_one = moments_pearson(S.linked_moments()) if rank == 0 else None
        comm.close()
        q.put((rank, r_all, m.n, r_one, moments_pearson(mx), mx.n, r_ex_one))
    finally:
        dist.destroy_process_group()


def test_nccl_moment_allreduce_two_gpus():
    """cfg4's collective through the C ABI: two ranks, one GPU each; the all-reduced r equals
    one GPU's r over the union sample range to 1e-12, for the sampler and the exhaustive pass."""
    n = C.c_int(0)
    _lib.lib().st_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    (_, r0, n0, r_one, x0, xn0, r_ex_one), (_, r1, n1, _, x1, xn1, _) = got
    assert r0 == r1 and n0 == n1 == 1_000_000
    assert abs(r0 - r_one) <= 1e-12
    assert x0 == x1 and xn0 == xn1 == 2500 * 2499 // 2 and abs(x0 - r_ex_one) <= 1e-12


# ------------------------------------------- kernel variants (round 2) -------
def _with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("shape", ["yule", "left_comb", "right_comb", "balanced"])
@pytest.mark.parametrize("block_shift", [0, 3])
def test_paired_record_kernel_matches_plain_kernel_and_oracle(shape, block_shift):
    """The paired-record pair kernel (one 32-byte load per endpoint; rd[mrca] from an endpoint
    or its sector neighbour) against the plain kernel and the oracle: leaf pairs, any-node
    pairs, repeated ids, MRCA ids, on both comb orientations (the record shift is voted per
    tree) and with tiny RMQ blocks so that every route is taken."""
    if shape == "yule":
        ft = synth.yule_tree(3000, seed=3)
    elif shape == "balanced":
        ft = synth.balanced_tree(2048, seed=3)
    else:
        ft = synth.caterpillar_tree(1500, seed=3)
        if shape == "right_comb":
            n = ft.size
            src = ft
            ft = synth.caterpillar_tree(1500, seed=3)
            rev = lambda a: np.where(a >= 0, n - 1 - a, -1).astype(np.int32)[::-1].copy()  # noqa: E731
            ft.parent, ft.left, ft.right = rev(src.parent), rev(src.right), rev(src.left)
            ft.distance = src.distance[::-1].copy()
            ft.root = n - 1 - src.root
    T = SuchTree.from_flat(ft, _block_shift=block_shift)
    assert T.index_info["layout"] == 1
    ot = O.OracleTree(ft.parent, ft.distance)
    rng = np.random.default_rng(8)
    p = np.concatenate([2 * rng.integers(0, ft.n_leaves, size=(300_001, 2)), rng.integers(0, ft.size, size=(300_000, 2)),
                        np.repeat(rng.integers(0, ft.size, size=(1000, 1)), 2, axis=1)]).astype(np.int64)
    want, wm = ot.distances_f64_climb(p, with_mrca=True)
    # the three compact pair kernels: paired records (lean), plain lean, plain generic
    for paired, lean in (("1", "1"), ("0", "1"), ("0", "0")):
        def both():
            return _with_env("SUCHTREE_B200_LEAN", lean, lambda: (T.distances_bulk(p), T.common_ancestors_bulk(p)))
        d, m = _with_env("SUCHTREE_B200_PAIRED", paired, both)
        assert np.array_equal(d, want), (paired, lean)
        assert np.array_equal(m, wm), (paired, lean)


@pytest.mark.parametrize("qpt", ["1", "2"])
def test_depth_only_quartet_kernel_variants(qpt):
    """1 or 2 quartets per thread, int64 and int32 ids on the device, ragged counts, any nodes
    and repeated ids: rows bit-exact against the oracle's literal restatement
    (MuchTree.pyx:1331-1376)."""
    import torch

    ft = synth.yule_tree(5000, seed=9)
    ot = O.OracleTree(ft.parent, ft.distance)
    rng = np.random.default_rng(10)
    for bs in (0, 2, "wide"):
        T = SuchTree.from_flat(ft, _wide=True) if bs == "wide" else SuchTree.from_flat(ft, _block_shift=bs)
        assert T.index_info["layout"] == (0 if bs == "wide" else 1)
        for n in (1, 2, 3, 255, 256, 257, 383, 384, 385, 100_003):
            q = np.concatenate([2 * rng.integers(0, ft.n_leaves, size=(n, 4)),
                                rng.integers(0, ft.size, size=(n, 4))]).astype(np.int64)
            q[::5, 1] = q[::5, 0]
            q[::9, 3] = q[::9, 1]
            q[::17] = q[::17, :1]  # all four equal
            want = ot.quartet_topologies(q)
            got = _with_env("SUCHTREE_B200_QPT", qpt, lambda: T.quartet_topologies_bulk(q))
            assert np.array_equal(got, want), (bs, n)
            dq = torch.from_numpy(q.astype(np.int32)).cuda()
            do = torch.empty_like(dq)
            _with_env("SUCHTREE_B200_QPT", qpt,
                      lambda: T.quartet_topologies_device(dq.data_ptr(), q.shape[0], do.data_ptr(), idx_bits=32))
            torch.cuda.synchronize()
            assert np.array_equal(do.cpu().numpy().astype(np.int64), want), (bs, n, "int32")
        bad = torch.tensor([[0, 2, 4, ft.size]], dtype=torch.int32).cuda()
        ob = torch.empty_like(bad)
        T.quartet_topologies_device(bad.data_ptr(), 1, ob.data_ptr(), idx_bits=32)
        with pytest.raises(InvalidNodeError):
            T.check_range()
        assert ob.cpu().numpy().tolist() == [[-1, -1, -1, -1]]


def test_more_callers_than_lanes_wait_their_turn(tree):
    """Twelve threads, four lanes per device: the extra callers block until a lane is free;
    every call returns its own correct result (distances, MRCA ids, quartets mixed)."""
    T, ot, ft = tree
    out, errs = {}, []

    def run(k):
        try:
            rng = np.random.default_rng(100 + k)
            for rep in range(4):
                n = int(rng.integers(1, 400_000))
                p = rng.integers(0, ft.size, size=(n, 2)).astype(np.int64)
                if k % 3 == 0:
                    got, want = T.distances_bulk(p), ot.distances_f64_climb(p)
                elif k % 3 == 1:
                    got, want = T.common_ancestors_bulk(p), ot.distances_f64_climb(p, with_mrca=True)[1]
                else:
                    q = rng.integers(0, ft.size, size=(n // 4 + 1, 4)).astype(np.int64)
                    got, want = T.quartet_topologies_bulk(q), ot.quartet_topologies(q)
                if not np.array_equal(got, want):
                    errs.append((k, rep, "mismatch"))
            out[k] = True
        except Exception as e:  # noqa: BLE001
            errs.append((k, repr(e)))

    th = [threading.Thread(target=run, args=(k,)) for k in range(12)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    assert len(out) == 12


# ------------------------------------------------------- by-name entry points, C glue -------
def test_by_name_entry_points_same_through_c_glue_and_python_walk():
    """distances_by_name / quartet_topologies_by_name (MuchTree.pyx:945-979, :1378-1422): the C
    name -> id walk and the Python one give the same lists, and the same errors."""
    from suchtree_b200.exceptions import NodeNotFoundError

    T = SuchTree(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data", "fishworm_guest.tree"))
    names = T.leaf_names
    rng = np.random.default_rng(8)
    pairs = [(names[a], names[b]) for a, b in rng.integers(0, len(names), size=(20000, 2))]
    quartets = [tuple(names[i] for i in rng.choice(len(names), 4, replace=False)) for _ in range(2000)]
    if _lib.py_glue() is None:
        pytest.skip("C glue not built here")
    d_c, q_c = T.distances_by_name(pairs), T.quartet_topologies_by_name(quartets)
    assert isinstance(d_c, list) and isinstance(d_c[0], float) and isinstance(q_c[0], frozenset)
    saved = _lib._py_glue
    _lib._py_glue = None
    try:
        d_p, q_p = T.distances_by_name(pairs), T.quartet_topologies_by_name(quartets)
    finally:
        _lib._py_glue = saved
    assert d_c == d_p and q_c == q_p
    ids = np.array([[T.leaves[a], T.leaves[b]] for a, b in pairs], dtype=np.int64)
    assert d_c == T.distances_bulk(ids).tolist()
    with pytest.raises(NodeNotFoundError):
        T.distances_by_name(pairs[:10] + [(names[0], "nope")])
    with pytest.raises(TypeError):
        T.distances_by_name(pairs[:10] + [(names[0], 3)])
    assert T.distances_by_name([[names[0], names[1]]]) == T.distances_by_name([(names[0], names[1])])


# ------------------------------------------------------------ joined link records -----------
def _moments(m):
    return (m.n, m.sx, m.sy, m.sxx, m.syy, m.sxy)


def test_joined_link_records_give_the_same_bits_as_separate_rows():
    """The handle's one-sector-per-link records (both trees' root distances and block keys, compact
    layouts only) feed the Philox sampler, the exhaustive moments and the clade scan; every sum must
    equal the separate-rows path bit for bit, and both must agree with the oracle's distances
    (MuchTree.pyx:2900-2934 pair order; :62-79 moments)."""
    n = 3000
    fa, fb = synth.yule_tree(n, seed=31), synth.yule_tree(n, seed=32)
    A, B = SuchTree.from_flat(fa), SuchTree.from_flat(fb)
    rng = np.random.default_rng(33)
    L = 700
    ll = np.stack([2 * rng.integers(0, n, L), 2 * rng.integers(0, n, L)], axis=1).astype(np.int64)
    got = {}
    for mode in ("0", "1"):
        os.environ["SUCHTREE_B200_JOINED"] = mode
        try:
            S = SuchLinkedTrees.from_linklist(A, B, ll)
            got[mode] = (_moments(S.sample_moments(200001, seed=5, first_sample=3, x0=1.0, y0=2.0)),
                         _moments(S.linked_moments(17, 200000, x0=1.0, y0=2.0)),
                         S.clade_pearson(min_links=3, max_links=400))
        finally:
            del os.environ["SUCHTREE_B200_JOINED"]
    assert got["0"][0] == got["1"][0] and got["0"][1] == got["1"][1]
    for key in ("r", "n_pairs", "n_links"):
        assert np.array_equal(got["0"][2][key], got["1"][2][key], equal_nan=True), key
    # against the oracle: exhaustive moments over the first pairs of the enumeration
    S = SuchLinkedTrees.from_linklist(A, B, ll)
    oa, ob = O.OracleTree(fa.parent, fa.distance), O.OracleTree(fb.parent, fb.distance)
    lk = S.linklist
    i, j = np.array([(i, j) for i in range(1, 60) for j in range(i)]).T
    da = oa.distances_f64(np.stack([lk[j, 1], lk[i, 1]], axis=1))
    db = ob.distances_f64(np.stack([lk[j, 0], lk[i, 0]], axis=1))
    m = S.linked_moments(0, len(i))
    assert m.n == len(i)
    assert abs(m.sx - da.sum()) <= 1e-9 * da.sum() and abs(m.sxy - (da * db).sum()) <= 1e-9 * (da * db).sum()


def test_joined_link_records_fall_back_when_a_tree_does_not_fit():
    """A caterpillar's block keys (depth up to 10^5) do not fit the packed word: the handle keeps
    separate rows, and the results are those of the generic path."""
    fa, fb = synth.caterpillar_tree(40000, seed=1), synth.yule_tree(40000, seed=2)
    A, B = SuchTree.from_flat(fa), SuchTree.from_flat(fb)
    rng = np.random.default_rng(3)
    ll = np.stack([2 * rng.integers(0, 40000, 500), 2 * rng.integers(0, 40000, 500)], axis=1).astype(np.int64)
    S = SuchLinkedTrees.from_linklist(A, B, ll)
    m1 = _moments(S.sample_moments(50000, seed=1))
    os.environ["SUCHTREE_B200_JOINED"] = "0"
    try:
        m0 = _moments(SuchLinkedTrees.from_linklist(A, B, ll).sample_moments(50000, seed=1))
    finally:
        del os.environ["SUCHTREE_B200_JOINED"]
    assert m0 == m1
    d = S.linked_distances()
    oa = O.OracleTree(fa.parent, fa.distance)
    assert np.array_equal(d["TreeA"][:2000], oa.distances_f64(d["ids_A"][:2000]))


# ----------------------------------------- compact arithmetic where rounding happens --------
def test_compact_distance_arithmetic_rounds_once_on_inexact_sums():
    """Edges 16 or 2^k, k in [-46, -40], depth 8: every root distance is exact in fp64 (compact
    layout), but rd[a] + rd[b] - 2 rd[mrca] needs up to 55 bits (about one pair in six).  The lean kernels (st_patristic_c: the
    double-double sum with its no-op steps removed), the generic kernel (full double-double) and
    the joined-record linked path must all return the correctly rounded value of the exact sum."""
    ft = synth.balanced_tree(256, seed=5)
    rng = np.random.default_rng(6)
    k = np.where(rng.random(ft.size) < 0.5, 4, rng.integers(-46, -39, size=ft.size))
    ft.distance = np.ldexp(1.0, k).astype(np.float32)
    ft.distance[ft.root] = -1.0
    T = SuchTree.from_flat(ft)
    assert T.index_info["layout"] == 1
    unit = 46
    units = [0] * ft.size  # exact root distances in units of 2^-46
    depth_first = list(T.traverse_preorder())
    for v in depth_first:
        p = int(ft.parent[v])
        units[v] = 0 if p == -1 else units[p] + (1 << (int(k[v]) + unit))
    p = rng.integers(0, ft.size, size=(200_000, 2)).astype(np.int64)
    m = T.common_ancestors_bulk(p)
    exact = [units[a] + units[b] - 2 * units[c] for (a, b), c in zip(p.tolist(), m.tolist())]
    inexact = sum(1 for d in exact if d and d.bit_length() - ((d & -d).bit_length() - 1) > 53)
    assert inexact > 10000  # the case under test does occur
    want = np.array([d / float(1 << unit) for d in exact])  # int / float: correctly rounded
    for paired, lean in (("1", "1"), ("0", "1"), ("0", "0")):
        d = _with_env("SUCHTREE_B200_PAIRED", paired, lambda: _with_env("SUCHTREE_B200_LEAN", lean, lambda: T.distances_bulk(p)))
        assert np.array_equal(d, want), (paired, lean)
    W = SuchTree.from_flat(ft, _wide=True)  # 32-byte records, double-double root distances
    assert np.array_equal(W.distances_bulk(p), want)
    D = T.pairwise_distances(list(range(0, ft.size, 3)))
    ids = np.arange(0, ft.size, 3)
    a, b = np.meshgrid(ids, ids, indexing="ij")
    assert np.array_equal(D.ravel(), T.distances_bulk(np.stack([a.ravel(), b.ravel()], axis=1)))
    # linked paths: joined records and separate rows
    leaves = T.leaf_node_ids
    ll = np.stack([rng.choice(leaves, 300), rng.choice(leaves, 300)], axis=1).astype(np.int64)
    for mode in ("0", "1"):
        S = _with_env("SUCHTREE_B200_JOINED", mode, lambda: SuchLinkedTrees.from_linklist(T, T, ll))
        res = _with_env("SUCHTREE_B200_JOINED", mode, lambda: S.linked_distances())
        ia = res["ids_A"]
        ma = T.common_ancestors_bulk(ia)
        wa = np.array([(units[x] + units[y] - 2 * units[c]) / float(1 << unit) for (x, y), c in zip(ia.tolist(), ma.tolist())])
        assert np.array_equal(res["TreeA"], wa)
        mo = _with_env("SUCHTREE_B200_JOINED", mode, lambda: S.linked_moments())
        assert abs(mo.sx - wa.sum()) <= 1e-12 * wa.sum()
