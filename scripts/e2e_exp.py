"""e2e (host-buffer) throughput of SuchTree.distances_bulk vs the pack fraction (run under gpurun)."""
import os, sys, time, subprocess
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    os.environ['SUCHTREE_B200_PACK_FRACTION'] = sys.argv[1]
    if sys.argv[1] == 'direct':
        os.environ['SUCHTREE_B200_HOST_PATH'] = 'direct'
    import numpy as np, torch
    from suchtree_b200 import SuchTree, synth, philox_host
    T = SuchTree.from_flat(synth.yule_tree(100000, seed=1))
    n = 100_000_000
    hp = torch.empty((n, 2), dtype=torch.int64).pin_memory()
    hp.copy_(torch.from_numpy(2 * np.random.default_rng(0).integers(0, 100000, size=(n, 2))))
    ho = torch.empty(n, dtype=torch.float64).pin_memory()
    P, Oo = hp.numpy(), ho.numpy()
    T.distances_bulk(P, out=Oo)
    best = 0
    for _ in range(5):
        t0 = time.perf_counter(); T.distances_bulk(P, out=Oo); dt = time.perf_counter() - t0
        best = max(best, n / dt)
    # pageable in/out as a plain user would pass them
    Pp = np.array(P[:20_000_000]); t0 = time.perf_counter(); r = T.distances_bulk(Pp); dt = time.perf_counter() - t0
    ok = np.array_equal(r, Oo[:20_000_000])
    print('fraction', sys.argv[1], 'pinned best %.3e pairs/s' % best, ' pageable %.3e' % (20_000_000 / dt), ok, flush=True)
else:
    for f in ('direct', '0', '0.3', '0.4', '0.5', '0.6', '0.75', '1'):
        subprocess.run([sys.executable, __file__, f])
