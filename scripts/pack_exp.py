import ctypes as C, os, sys, subprocess
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    os.environ['SUCHTREE_B200_HOST_THREADS'] = sys.argv[1]
    from suchtree_b200 import _lib
    v = C.c_double()
    for n in (4 << 20, 32 << 20):
        rc = _lib.lib().st_bench_pack(n, 10, C.byref(v))
        print('threads', sys.argv[1], 'pairs', n, rc, '%.3e pairs/s  %.1f GB/s moved' % (v.value, v.value * 24 / 1e9), flush=True)
else:
    for t in (1, 2, 4, 8, 16):
        subprocess.run([sys.executable, __file__, str(t)])
