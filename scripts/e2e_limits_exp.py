"""What the host interface of one GPU carries in each direction and both at once
(st_bench_copy, pinned memory, 64 MiB chunks); run under gpurun."""
import ctypes as C, json, sys
sys.path.insert(0, '.')
from suchtree_b200 import _lib
L = _lib.lib()
n = 100_000_000
out = {}
for name, h2d, d2h in (("h2d_only_16B", 16, 0), ("d2h_only_8B", 0, 8), ("h2d_8B_d2h_8B", 8, 8),
                       ("h2d_16B_d2h_8B", 16, 8), ("h2d_12B_d2h_8B", 12, 8), ("h2d_5B_d2h_8B", 5, 8)):
    sec = C.c_double()
    rc = L.st_bench_copy(0, h2d * n, d2h * n, 64 << 20, 5, C.byref(sec))
    out[name] = {"rc": rc, "s": sec.value, "h2d_GBs": h2d * n / sec.value / 1e9, "d2h_GBs": d2h * n / sec.value / 1e9,
                 "pairs_per_s": n / sec.value}
    print(name, out[name], flush=True)
json.dump(out, open("gpurun_out/e2e_limits.json", "w"), indent=1)
