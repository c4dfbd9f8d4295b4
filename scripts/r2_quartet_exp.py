"""Quartet kernel variants (run under gpurun): resident CTAs asked of the compiler, prefetch,
quartets per thread, int64 / int32 ids.  -> gpurun_out/r2_quartet_exp.json"""
import json, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import oracle as O
from suchtree_b200 import SuchTree, synth

dev = torch.device('cuda', 0)
stream = torch.cuda.current_stream(dev)
sptr = stream.cuda_stream


def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps


ft = synth.yule_tree(100000, seed=1)
T = SuchTree.from_flat(ft, device=0)
ot = O.OracleTree(ft.parent, ft.distance)
nq = 50_000_000
g = torch.Generator(device=dev).manual_seed(21)
q64 = 2 * torch.randint(0, 100000, (nq, 4), generator=g, device=dev, dtype=torch.int64)
o64 = torch.empty_like(q64)
q32 = q64.to(torch.int32)
o32 = torch.empty_like(q32)
want = ot.quartet_topologies(q64[:300000].cpu().numpy())
res = {}
for name, env in (('default_qpt1_256x4', {}), ('qpt1_384x3', {'SUCHTREE_B200_QT': '384'}), ('qpt2', {'SUCHTREE_B200_QPT': '2'})):
    for k in ('SUCHTREE_B200_QT', 'SUCHTREE_B200_QPT'):
        os.environ.pop(k, None)
    os.environ.update(env)
    s64 = timed(lambda: T.quartet_topologies_device(q64.data_ptr(), nq, o64.data_ptr(), stream=sptr))
    s32 = timed(lambda: T.quartet_topologies_device(q32.data_ptr(), nq, o32.data_ptr(), stream=sptr, idx_bits=32))
    res[name] = {'int64_quartets_per_s': nq / s64, 'int32_quartets_per_s': nq / s32,
                 'equals_oracle': bool(np.array_equal(o64[:300000].cpu().numpy(), want)),
                 'int32_equals_int64': bool(torch.equal(o32.to(torch.int64), o64))}
    print(name, res[name], flush=True)
T.check_range(sptr)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r2_quartet_exp.json', 'w'), indent=1)
