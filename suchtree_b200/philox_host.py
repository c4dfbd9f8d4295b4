"""numpy restatement of Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011 -- the
Random123 generator) and of the index mapping the kernels use, for checking the
device sampler / pair generator on an identical index stream.  Host-side helper (tests, bench sample generation); the device code is in csrc/st_device.cuh."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint64 array (low 64 bits of the 128-bit counter, high bits 0);
    key: python int (64 bit).  Returns four uint32 arrays."""
    counter = np.asarray(counter, dtype=np.uint64)
    c0 = counter & MASK
    c1 = counter >> np.uint64(32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def bounded(u, n):
    """floor(u * n / 2^32), the kernels' __umulhi mapping."""
    return ((u.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def random_leaf_pairs(n_leaves, seed, first_pair, n):
    """What st_random_leaf_pairs_device writes (int64 [n,2] of leaf ids)."""
    assert first_pair % 2 == 0
    n2 = (n + 1) // 2
    x, y, z, w = philox4x32_10(np.arange(first_pair // 2, first_pair // 2 + n2, dtype=np.uint64), seed)
    out = np.empty((2 * n2, 2), dtype=np.int64)
    out[0::2, 0] = 2 * bounded(x, n_leaves)
    out[0::2, 1] = 2 * bounded(y, n_leaves)
    out[1::2, 0] = 2 * bounded(z, n_leaves)
    out[1::2, 1] = 2 * bounded(w, n_leaves)
    return out[:n]


def sampled_links(n_links, seed, first_sample, n):
    """(l1, l2) link indices of samples [first_sample, first_sample+n) as
    st_sample_moments draws them: sample s uses words (2h, 2h+1) of
    Philox(counter = s >> 1), h = s & 1."""
    s = np.arange(first_sample, first_sample + n, dtype=np.uint64)
    x, y, z, w = philox4x32_10(s >> np.uint64(1), seed)
    odd = (s & np.uint64(1)).astype(bool)
    l1 = np.where(odd, bounded(z, n_links), bounded(x, n_links))
    l2 = np.where(odd, bounded(w, n_links), bounded(y, n_links))
    return l1, l2
