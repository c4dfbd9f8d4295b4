"""pearson(x, y) on host vectors: pageable and page-locked inputs (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from suchtree_b200 import pearson, _lib
rng = np.random.default_rng(0)
for n in (1000, 1_000_000, 100_000_000):
    x = rng.random(n); y = 0.5 * x + rng.random(n)
    want = float(np.corrcoef(x, y)[0, 1]) if n <= 1_000_000 else None
    r = pearson(x, y)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); r = pearson(x, y); ts.append(time.perf_counter() - t0)
    px = _lib.pinned_empty((n,), np.float64); py = _lib.pinned_empty((n,), np.float64)
    px[:] = x; py[:] = y
    rp = pearson(px, py)
    tp = []
    for _ in range(5):
        t0 = time.perf_counter(); rp = pearson(px, py); tp.append(time.perf_counter() - t0)
    print(n, "pageable %.3f ms (%.1f GB/s)  pinned %.3f ms (%.1f GB/s)  r %.12f %.12f %s" % (
        1e3 * min(ts), 16 * n / min(ts) / 1e9, 1e3 * min(tp), 16 * n / min(tp) / 1e9, r, rp, want), flush=True)
