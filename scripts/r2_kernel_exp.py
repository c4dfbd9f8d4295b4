"""Round-2 kernel experiments (run under gpurun): paired-record pair kernel vs the plain one on
Yule / balanced / both caterpillar orientations, and the depth-only quartet kernel (1 or 2
quartets per thread, int64 or int32 ids).  -> gpurun_out/r2_kernel_exp.json"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import oracle as O
from suchtree_b200 import SuchTree, _lib, synth

dev = torch.device('cuda', 0)
stream = torch.cuda.current_stream(dev)
sptr = stream.cuda_stream
res = {}


def timed(fn, steps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps


def mirrored(ft):
    """right comb: mirror image of synth.caterpillar_tree (leaf k's parent is id + 1)"""
    n = ft.size
    m = synth.caterpillar_tree(ft.n_leaves, seed=3)
    rev = lambda a: np.where(a >= 0, n - 1 - a, -1).astype(np.int32)[::-1].copy()
    m.parent, l, r = rev(ft.parent), rev(ft.right), rev(ft.left)
    m.left, m.right = l, r
    m.distance = ft.distance[::-1].copy()
    m.root = n - 1 - ft.root
    return m


trees = {'yule100k': synth.yule_tree(100000, seed=1), 'balanced1M': synth.balanced_tree(1_000_000, seed=3),
         'caterpillar1M_left': synth.caterpillar_tree(1_000_000, seed=3)}
trees['caterpillar1M_right'] = mirrored(trees['caterpillar1M_left'])
n = 400_000_000
pairs = torch.empty((n, 2), dtype=torch.int32, device=dev)
out = torch.empty(n, dtype=torch.float64, device=dev)
ref = torch.empty(n, dtype=torch.float64, device=dev)
for name, ft in trees.items():
    T = SuchTree.from_flat(ft, device=0)
    T.random_leaf_pairs_device(3, 0, n, pairs.data_ptr(), idx_bits=32, stream=sptr)
    entry = {'layout': int(T.index_info['layout'])}
    for paired in ('0', '1'):
        os.environ['SUCHTREE_B200_PAIRED'] = paired
        buf = ref if paired == '0' else out
        sec = timed(lambda: T.distances_device(pairs.data_ptr(), n, buf.data_ptr(), idx_bits=32, stream=sptr))
        entry['pairs_per_s_paired' + paired] = n / sec
    entry['paired_equals_plain'] = bool(torch.equal(out, ref))
    # oracle spot check
    hp = pairs[:200000].cpu().numpy().astype(np.int64)
    ot = O.OracleTree(ft.parent, ft.distance)
    want, wm = ot.distances_f64_climb(hp, with_mrca=True)
    entry['plain_equals_oracle'] = bool(np.array_equal(ref[:200000].cpu().numpy(), want))
    mr = torch.empty(200000, dtype=torch.int32, device=dev)
    T.distances_device(pairs.data_ptr(), 200000, out.data_ptr(), idx_bits=32, d_mrca_ptr=mr.data_ptr(), stream=sptr)
    torch.cuda.synchronize()
    entry['paired_mrca_equals_oracle'] = bool(np.array_equal(mr.cpu().numpy(), wm)) and bool(
        np.array_equal(out[:200000].cpu().numpy(), want))
    # int64 ids + any-node pairs through the paired kernel
    g = torch.Generator(device=dev).manual_seed(5)
    anyp = torch.randint(0, ft.size, (1_000_001, 2), generator=g, device=dev, dtype=torch.int64)
    o2 = torch.empty(1_000_001, dtype=torch.float64, device=dev)
    T.distances_device(anyp.data_ptr(), 1_000_001, o2.data_ptr(), idx_bits=64, stream=sptr)
    torch.cuda.synchronize()
    sel = anyp[:100000].cpu().numpy()
    entry['paired_any_nodes_int64_equals_oracle'] = bool(np.array_equal(o2[:100000].cpu().numpy(), ot.distances_f64_climb(sel)))
    T.check_range(sptr)
    res[name] = entry
    print(name, entry, flush=True)
    del T
del pairs, out, ref
os.environ.pop('SUCHTREE_B200_PAIRED', None)

# ---- quartets
ft = trees['yule100k']
T = SuchTree.from_flat(ft, device=0)
ot = O.OracleTree(ft.parent, ft.distance)
nq = 50_000_000
g = torch.Generator(device=dev).manual_seed(21)
q64 = 2 * torch.randint(0, 100000, (nq, 4), generator=g, device=dev, dtype=torch.int64)
o64 = torch.empty_like(q64)
q32 = q64.to(torch.int32)
o32 = torch.empty_like(q32)
entry = {}
for qpt in ('1', '2'):
    os.environ['SUCHTREE_B200_QPT'] = qpt
    s64 = timed(lambda: T.quartet_topologies_device(q64.data_ptr(), nq, o64.data_ptr(), stream=sptr))
    s32 = timed(lambda: T.quartet_topologies_device(q32.data_ptr(), nq, o32.data_ptr(), stream=sptr, idx_bits=32))
    entry['qpt%s_int64_quartets_per_s' % qpt] = nq / s64
    entry['qpt%s_int32_quartets_per_s' % qpt] = nq / s32
    entry['qpt%s_int32_equals_int64' % qpt] = bool(torch.equal(o32.to(torch.int64), o64))
    hq = q64[:300000].cpu().numpy()
    entry['qpt%s_equals_oracle' % qpt] = bool(np.array_equal(o64[:300000].cpu().numpy(), ot.quartet_topologies(hq)))
    # any nodes, repeated ids
    rng = np.random.default_rng(7)
    anyq = rng.integers(0, ft.size, size=(200000, 4)).astype(np.int64)
    anyq[::7, 1] = anyq[::7, 0]; anyq[::11, 3] = anyq[::11, 2]; anyq[::13, 2] = anyq[::13, 0]
    entry['qpt%s_any_nodes_equals_oracle' % qpt] = bool(np.array_equal(T.quartet_topologies_bulk(anyq), ot.quartet_topologies(anyq)))
T.check_range(sptr)
res['quartets_yule100k'] = entry
print(entry, flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r2_kernel_exp.json', 'w'), indent=1)
