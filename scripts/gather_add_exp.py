"""LSU + TMA gather4 additivity (run under gpurun): 16 LSU warps per CTA plus TW TMA warps
doing pct % as many rounds.  Prints total sectors/s."""
import ctypes as C, os, sys
sys.path.insert(0, '.')
from suchtree_b200 import _lib
L = _lib.lib()
B = _lib.bench_lib()
for tw in (0, 2, 4, 8, 12, 16):
    for pct in ((0,) if tw == 0 else (25, 50, 75, 100, 150)):
        os.environ['SUCHTREE_B200_GATHER_MODE'] = 'add%d_%d' % (tw, pct)
        v = C.c_double()
        rc = B.st_bench_gather(0, int(6.4e6), 256, 5, C.byref(v))
        print('tma_warps %2d  tma_rounds %3d %%  rc %d  %.3e sectors/s' % (tw, pct, rc, v.value), _lib.last_error() if rc else '', flush=True)
