"""Lean compact pair kernel: launch shapes and pairs per thread (run under gpurun).
-> gpurun_out/r2_lean_exp.json"""
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


res = {}
n = 200_000_000
pairs = torch.empty((n, 2), dtype=torch.int32, device=dev)
out = torch.empty(n, dtype=torch.float64, device=dev)
ref = torch.empty(n, dtype=torch.float64, device=dev)
mr = torch.empty(n, dtype=torch.int32, device=dev)
mref = torch.empty(n, dtype=torch.int32, device=dev)
for name, ft in (('yule100k', synth.yule_tree(100000, seed=1)), ('balanced1M', synth.balanced_tree(1_000_000, seed=3)),
                 ('caterpillar1M', synth.caterpillar_tree(1_000_000, seed=3))):
    T = SuchTree.from_flat(ft, device=0)
    T.random_leaf_pairs_device(3, 0, n, pairs.data_ptr(), idx_bits=32, stream=sptr)
    os.environ['SUCHTREE_B200_PAIRED'] = '0'
    entry = {}
    for label, env in (('generic_p2_512x2', {'SUCHTREE_B200_LEAN': '0'}), ('lean_p2_512x2', {'SUCHTREE_B200_LEAN': '1'}),
                       ('paired_generic', {'SUCHTREE_B200_LEAN': '0', 'SUCHTREE_B200_PAIRED': '1'}),
                       ('paired_lean', {'SUCHTREE_B200_LEAN': '1', 'SUCHTREE_B200_PAIRED': '1'}),
                       ('lean_p4_256x3', {'SUCHTREE_B200_LEAN': '1', 'SUCHTREE_B200_PPT': '4'})):
        os.environ['SUCHTREE_B200_PAIRED'] = '0'
        for k in ('SUCHTREE_B200_LEAN', 'SUCHTREE_B200_PPT'):
            os.environ.pop(k, None)
        os.environ.update(env)
        buf = ref if label == 'generic_p2_512x2' else out
        sec = timed(lambda: T.distances_device(pairs.data_ptr(), n, buf.data_ptr(), idx_bits=32, stream=sptr))
        mb = mref if label == 'generic_p2_512x2' else mr
        T.distances_device(pairs.data_ptr(), n, buf.data_ptr(), idx_bits=32, d_mrca_ptr=mb.data_ptr(), stream=sptr)
        torch.cuda.synchronize()
        entry[label] = {'pairs_per_s': n / sec}
        if buf is out:
            entry[label]['equals_generic'] = bool(torch.equal(out, ref)) and bool(torch.equal(mr, mref))
    # any-node int64 pairs incl. repeats through the default lean path vs the oracle
    for k in ('SUCHTREE_B200_LEAN', 'SUCHTREE_B200_PPT'):
        os.environ.pop(k, None)
    rng = np.random.default_rng(5)
    p = rng.integers(0, ft.size, size=(300_001, 2)).astype(np.int64)
    p[::7, 1] = p[::7, 0]
    ot = O.OracleTree(ft.parent, ft.distance)
    want, wm = ot.distances_f64_climb(p, with_mrca=True)
    entry['lean_any_nodes_equals_oracle'] = bool(np.array_equal(T.distances_bulk(p), want)) and bool(
        np.array_equal(T.common_ancestors_bulk(p), wm))
    T.check_range(sptr)
    res[name] = entry
    print(name, json.dumps(entry), flush=True)
    del T
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r2_lean_exp.json', 'w'), indent=1)
