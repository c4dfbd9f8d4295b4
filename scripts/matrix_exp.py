"""Matrix writer timing vs a plain 10 GB device fill (run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
import torch
from suchtree_b200 import SuchTree, synth, _lib
dev = torch.device('cuda', 0)
ft = synth.yule_tree(100_000, seed=1)
T = SuchTree.from_flat(ft, device=0)
n = 100_000
rows = 12544
block = torch.empty((rows, n), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev)
sptr = stream.cuda_stream
def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
def mat():
    _lib.check(_lib.lib().st_distance_matrix(T._handle, None, n, 0, rows, block.data_ptr(), 1, sptr))
ms = timed(mat)
print('matrix  %.3f ms  %.3e el/s  %.1f GB/s' % (ms, rows * n / ms * 1e3, rows * n * 8 / ms / 1e6))
ms = timed(lambda: block.fill_(1.0))
print('fill_   %.3f ms  %.1f GB/s' % (ms, rows * n * 8 / ms / 1e6))
ms = timed(lambda: block.zero_())
print('zero_   %.3f ms  %.1f GB/s' % (ms, rows * n * 8 / ms / 1e6))
# single call latency, synchronised
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); mat(); torch.cuda.synchronize()
    print('one call wall %.3f ms' % ((time.perf_counter() - t0) * 1e3))
