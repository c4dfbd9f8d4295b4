"""Which hardware path gathers random 32-byte sectors fastest?  (run under gpurun)
   python scripts/gather_exp.py [mode ...]  -> gpurun_out/gather_exp.json"""
import ctypes as C, os, sys, json
sys.path.insert(0, '.')
from suchtree_b200 import _lib
L = _lib.lib()
B = _lib.bench_lib()
res = {}
modes = sys.argv[1:] or ['lsu', 'tex', 'mix', 'bulk', 'g4', 'g4mix1', 'g4mix2', 'g4mix3']
for mode in modes:
    os.environ['SUCHTREE_B200_GATHER_MODE'] = mode
    for mb in (6.4, 64):
        for loads in (64, 256):
            v = C.c_double()
            rc = B.st_bench_gather(0, int(mb * 1e6), loads, 5, C.byref(v))
            res['%s_%gMB_%d' % (mode, mb, loads)] = (rc, v.value)
            print(mode, mb, loads, rc, '%.3e sectors/s' % v.value, _lib.last_error() if rc else '', flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/gather_exp.json', 'w'), indent=1)
