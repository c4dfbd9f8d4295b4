"""Rate of the host bit packer alone (no device work); run on the bench box."""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import _lib
L = _lib.lib()
n = 50_000_000
ids = 2 * np.random.default_rng(0).integers(0, 100000, size=(n, 2))
for w in (18, 21):
    out = np.empty((n * 2 * w + 63) // 64 + 1, dtype=np.uint64)
    acc = C.c_uint64()
    L.st_host_pack_pairs(ids.ctypes.data, 2, 1, n, w, out.ctypes.data, C.byref(acc))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); L.st_host_pack_pairs(ids.ctypes.data, 2, 1, n, w, out.ctypes.data, C.byref(acc)); ts.append(time.perf_counter() - t0)
    print("w", w, "%.3e pairs/s (%.1f GB/s read)" % (n / min(ts), 16 * n / min(ts) / 1e9), flush=True)
