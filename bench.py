#!/usr/bin/env python
"""Benchmark of the batched patristic-distance path (BASELINE.json metric:
patristic distance pairs/sec, cfg 2: 100k-leaf random binary tree, 1e8 random leaf
pairs per step through distances()).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Our arm (default): one process per GPU (torchrun for N>1, rank i -> cuda:i).  The
tree index is replicated, every rank owns its own 1e8-pair shard of the Philox
pair stream (weak scaling, no data-path collective).  A step = ONE launch of the
pair kernel over the rank's device-resident pairs.  `value` = pairs all ranks
processed / max-over-ranks device time (CUDA events on the launching stream).
`e2e` = the same metric through the drop-in call SuchTree.distances_bulk() on
PINNED HOST int64 pairs (H2D + kernel + D2H inside the timed region).
Rank 0 at N=1 also times the unmodified reference (oracle/_ref, its own Cython
machine code) on a bounded sample of the same pairs -> `cpu_baseline`.

Reference arm (--impl reference): the unmodified reference's distances_bulk() on
the box's host cores (fork pool over contiguous pair blocks, the decomposition the
reference's docs recommend), each step a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

TREE_LEAVES = 100_000
TREE_SEED = 1
PAIR_SEED = 2
PAIRS_PER_STEP = 100_000_000
METRIC = "patristic_distance_pairs_per_sec"
UNIT = "pairs/s"
WORKLOAD = ("cfg2: simulated 100k-leaf random binary (Yule, seed 1) tree, 1e8 random leaf-id "
            "pairs per GPU per step via distances()")


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- #
# reference (CPU) side -- the one place bench.py executes oracle/
# --------------------------------------------------------------------------- #
_REF_TREE = None
_REF_PAIRS = None


def _ref_worker(span):
    # the pair array and the tree are inherited through fork (copy-on-write), as in
    # the multiprocessing recipe of the reference's docs; only results travel back
    return _REF_TREE.distances_bulk(_REF_PAIRS[span[0]:span[1]])


class ReferenceCPU:
    """The unmodified reference extension (oracle/_ref) on the bench tree, or the C
    port (oracle/st_oracle.c, literal fp32 arithmetic) if the extension is absent."""

    def __init__(self, flat_tree):
        global _REF_TREE
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        self.kind = "port"
        mod = None
        try:
            import ref_loader

            mod = ref_loader.load_reference(build_if_missing=os.path.exists("/root/reference"))
        except Exception as e:  # pragma: no cover - diagnostics only
            print("[bench] reference extension unavailable: %r" % (e,), file=sys.stderr)
        from suchtree_b200 import synth

        if mod is not None:
            ft = flat_tree
            if ft.leaves is None:
                ft.leaves = {"L%d" % k: 2 * k for k in range(ft.n_leaves)}
            nwk = synth.to_newick(ft)
            t0 = time.time()
            _REF_TREE = mod.SuchTree(nwk)
            self.load_s = time.time() - t0
            assert _REF_TREE.size == ft.size
            self.kind = "reference"
        else:
            import oracle as O

            ot = O.OracleTree(flat_tree.parent, flat_tree.distance)

            class _Port:
                def distances_bulk(self, pairs):
                    return ot.distances_f32(pairs)

            _REF_TREE = _Port()
            self.load_s = 0.0

    def run(self, pairs, cores):
        """pairs/s of distances_bulk over `pairs` (int64 [n,2]) on `cores` processes."""
        if cores <= 1:
            t0 = time.perf_counter()
            out = _REF_TREE.distances_bulk(pairs)
            return pairs.shape[0] / (time.perf_counter() - t0), out
        import multiprocessing as mp

        global _REF_PAIRS
        _REF_PAIRS = pairs
        ctx = mp.get_context("fork")  # threads do not scale: the reference holds the GIL
        edges = np.linspace(0, pairs.shape[0], cores * 4 + 1).astype(np.int64)
        spans = list(zip(edges[:-1], edges[1:]))
        with ctx.Pool(cores) as pool:
            pool.map(_ref_worker, [(0, 1000)] * cores)  # warm the workers
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, spans)
            dt = time.perf_counter() - t0
        return pairs.shape[0] / dt, np.concatenate(res)


def sample_pairs_host(n_leaves, seed, first, n):
    """The bench's pair stream restated on the host (same Philox stream the device
    generator produces; suchtree_b200/philox_host.py, checked by tests/)."""
    from suchtree_b200 import philox_host

    return philox_host.random_leaf_pairs(n_leaves, seed, first, n)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from suchtree_b200 import synth

    ft = synth.yule_tree(TREE_LEAVES, seed=TREE_SEED)
    ref = ReferenceCPU(ft)
    cores = host_cores()
    # calibrate a bounded per-step sample: ~3 s of all-core work per step
    probe = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, 200_000)
    rate1, _ = ref.run(probe, 1)
    per_step = int(min(30_000_000, max(200_000, rate1 * cores * 0.6 * 3.0)))
    per_step -= per_step % 2
    pairs = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, per_step)
    for _ in range(max(args.warmup, 0) and 1):
        ref.run(pairs[: per_step // 4], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.run(pairs, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "tree_leaves": TREE_LEAVES, "pairs_per_step": per_step,
                   "note": "bounded sample of the 1e8-pair step; pool start-up outside the rate"},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": ref.kind, "cpu_model": cpu_model(),
            "sample": "%d pairs/step x %d steps of the cfg2 pair stream, fork Pool(%d) over "
                      "contiguous blocks; single core: %.3g pairs/s" % (per_step, args.steps, cores, rate1),
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler(threading.Thread):
    def __init__(self, device_index, period=0.004):  # the timed region is tens of milliseconds
        super().__init__(daemon=True)
        self.idx, self.period = device_index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get("k_pairs_dram_bytes_per_launch")
    except Exception:
        return None


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus, file=sys.stderr)
            return 2
    from suchtree_b200 import synth

    ft = synth.yule_tree(TREE_LEAVES, seed=TREE_SEED)
    n_pairs = args.pairs

    # ---- CPU baseline first (fork pool before this process touches CUDA)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = ReferenceCPU(ft)
        cores = host_cores()
        probe = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, 200_000)
        rate1, out1 = ref.run(probe, 1)
        n_s = int(min(n_pairs, max(400_000, rate1 * 4.0)))  # ~4 s on one core
        n_s -= n_s % 2
        sample = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, n_s)
        rate1, ref_out = ref.run(sample, 1)
        rate_all, _ = ref.run(sample, cores) if cores > 1 else (rate1, None)
        cpu_baseline = {
            "value": rate_all, "unit": UNIT, "cores": cores, "kind": ref.kind,
            "single_core_value": rate1, "cpu_model": cpu_model(), "os_cpu_count": os.cpu_count(),
            "sample": "first %d pairs of the step's Philox stream through the reference's "
                      "distances_bulk; 1 core, then fork Pool(%d)" % (n_s, cores),
        }
        cpu_sample = (sample, ref_out)
    else:
        cpu_sample = None

    import torch

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator
        # comes up (its version banner) goes to stderr -- fd 1 points at fd 2 meanwhile
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from suchtree_b200 import SuchTree, _lib

    T = SuchTree.from_flat(ft, device=local)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    pairs = torch.empty((n_pairs, 2), dtype=torch.int32, device=dev)
    out = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    # rank r owns pairs [r*n, (r+1)*n) of the global stream
    T.random_leaf_pairs_device(PAIR_SEED, rank * n_pairs, n_pairs, pairs.data_ptr(), idx_bits=32, stream=sptr)
    torch.cuda.synchronize()

    def step():
        T.distances_device(pairs.data_ptr(), n_pairs, out.data_ptr(), idx_bits=32, stream=sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    T.check_range(sptr)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_pairs * args.steps / (ms_max * 1e-3)

    # ---- parity spot check against the reference's own output (rank 0, N=1)
    parity = None
    if cpu_sample is not None:
        sample, ref_out = cpu_sample
        got = out[: sample.shape[0]].cpu().numpy()
        # reference accumulates in fp32 (MuchTree.pyx:924): depth-scaled tolerance
        parity = bool(np.all(np.abs(got - ref_out) <= 2e-7 * T.depth * np.maximum(got, 1e-30)))
        host_pairs = pairs[: sample.shape[0]].cpu().numpy().astype(np.int64)
        parity = parity and bool(np.array_equal(host_pairs, sample))

    # ---- e2e: the drop-in call on pinned host buffers
    n_e2e = min(n_pairs, args.e2e_pairs)
    h_pairs = torch.empty((n_e2e, 2), dtype=torch.int64).pin_memory()
    h_pairs.copy_(pairs[:n_e2e].to(torch.int64))
    h_out = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    np_pairs, np_out = h_pairs.numpy(), h_out.numpy()
    T.distances_bulk(np_pairs, out=np_out)  # warm-up: staging buffers
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        T.distances_bulk(np_pairs, out=np_out)
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e * e2e_steps / float(te.item())
    e2e_ok = bool(torch.equal(h_out.to(dev), out[:n_e2e]))

    # ---- rooflines (rank 0's kernel; all ranks run the same launch)
    peak, peak_src = measured_peak()
    per_launch_s = ms * 1e-3 / args.steps
    achieved = 16.0 * n_pairs / per_launch_s / 1e9
    gather = None
    if rank == 0:
        import ctypes as C

        sps = C.c_double(0)
        rc = _lib.lib().st_bench_gather(local, int(T.index_info["index_bytes"]), 64, 5, C.byref(sps))
        if rc == 0:
            gather = {
                "what": "random 32-byte-sector gathers over an index-sized buffer (%d bytes), measured by "
                        "st_bench_gather in this run" % T.index_info["index_bytes"],
                "peak_sectors_per_s": sps.value,
                "achieved_sectors_per_s": 2.0 * n_pairs / per_launch_s,
                "frac": 2.0 * n_pairs / per_launch_s / sps.value,
                "sectors_per_pair": 2,
                "note": "rec[lo] + rec[hi]; rd[mrca] comes from the shared-memory block table for ~97 % of "
                        "far-apart pairs (DESIGN.md 4.1)",
            }
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "tree_leaves": TREE_LEAVES, "tree_nodes": T.size, "tree_depth": T.depth,
            "pairs_per_step_per_gpu": n_pairs, "pair_dtype": "int32x2", "result_dtype": "f64",
            "parallelism": "index replicated, pair stream sharded per GPU, no collective",
            "l2": "inputs larger than L2: %.0f MB of pairs+results streamed per step vs 126 MB L2"
                  % (16.0 * n_pairs / 1e6),
            "index_bytes": int(T.index_info["index_bytes"]),
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "k_pairs<int32,VEC>",
            "algorithmic_bytes_per_pair": 16,
            "note": "the kernel is bound by the per-SM L1TEX rate of random sector gathers (2 per pair), see gather_roofline",
        },
        "gather_roofline": gather,
        "cpu_baseline": cpu_baseline,
        "e2e": {
            "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n_e2e, "d2h_bytes_per_step": 8 * n_e2e,
            "pairs_per_step_per_gpu": n_e2e, "steps": e2e_steps, "matches_device_path": e2e_ok,
            "call": "SuchTree.distances_bulk(int64 (n,2) pinned host array, out=pinned float64)",
        },
        "gpu_launches": args.steps,
        "clocks": clocks,
        "parity_vs_reference_sample": parity,
    }
    if not args.no_other_workloads:
        del pairs, out, h_pairs, h_out
        torch.cuda.empty_cache()
        try:
            line["other_workloads"] = run_other_workloads(args, rank, world, local, dev, peak)
        except Exception as e:  # the headline line must still be printed
            line["other_workloads"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



# --------------------------------------------------------------------------- #
# the other BASELINE.json configs (cfg 3, 4, 5), same process, short runs
# --------------------------------------------------------------------------- #
def _timed(stream, fn, steps, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps


def run_other_workloads(args, rank, world, local, dev, peak):
    """cfg3 (1M-leaf caterpillar + balanced, 1.25e9 pairs per GPU), cfg5 (100k x 100k
    matrix, 12,500-row block per GPU, device-resident) and cfg4 (two 100k-leaf trees,
    1.25e8 Philox-sampled link pairs per GPU, moments all-reduced).  Every rank runs
    its own shard; times are max over ranks; rank 0 reports."""
    import torch
    import torch.distributed as dist

    import ctypes as C

    from suchtree_b200 import SuchTree, _lib, shard, synth
    from suchtree_b200.linked import moments_pearson

    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}
    # ---- cfg1: the reference's own CPU-runnable case -- gopher tree (29 nodes), 1e6 random
    #      leaf pairs: GPU (device-resident and through the drop-in host call) beside the
    #      unmodified reference on one host core, results compared
    if rank == 0:
        G = SuchTree(os.path.join(REPO, "tests", "golden", "data", "test.tree"), device=local)
        p1 = (np.random.default_rng(0).integers(0, 15, size=(1_000_000, 2)) * 2).astype(np.int64)
        d_p1 = torch.from_numpy(p1).to(dev)
        d_o1 = torch.empty(p1.shape[0], dtype=torch.float64, device=dev)
        sec = _timed(stream, lambda: G.distances_device(d_p1.data_ptr(), p1.shape[0], d_o1.data_ptr(), idx_bits=64,
                                                        stream=sptr), steps=20)
        G.distances_bulk(p1)
        t0 = time.perf_counter()
        for _ in range(5):
            got = G.distances_bulk(p1)
        host_s = (time.perf_counter() - t0) / 5
        entry = {"workload": "cfg1: data/gopher-louse gopher tree (29 nodes), 1e6 random leaf-id pairs (int64)",
                 "pairs_per_s_device_resident": p1.shape[0] / sec, "pairs_per_s_host_call": p1.shape[0] / host_s,
                 "matches_device_path": bool(np.array_equal(got, d_o1.cpu().numpy()))}
        try:
            sys.path.insert(0, os.path.join(REPO, "oracle"))
            import ref_loader

            mod = ref_loader.load_reference(build_if_missing=False) if world == 1 else None
            if mod is not None:
                R = mod.SuchTree(os.path.join(REPO, "tests", "golden", "data", "test.tree"))
                R.distances_bulk(p1[:1000])
                t0 = time.perf_counter()
                want = R.distances_bulk(p1)
                entry["reference_pairs_per_s_one_core"] = p1.shape[0] / (time.perf_counter() - t0)
                # fp32 accumulation in the reference (MuchTree.pyx:924): depth-scaled tolerance
                entry["parity_vs_reference"] = bool(np.all(np.abs(got - want) <= 2e-7 * R.depth * np.maximum(got, 1e-30)))
        except Exception as e:  # diagnostics only
            entry["reference_error"] = repr(e)[:80]
        res["cfg1_gopher"] = entry
        del G

    # ---- cfg3: deep-path worst case + balanced, 1M leaves
    n3 = args.cfg3_pairs
    pairs = torch.empty((n3, 2), dtype=torch.int32, device=dev)
    out = torch.empty(n3, dtype=torch.float64, device=dev)
    for shape, gen in (("caterpillar", synth.caterpillar_tree), ("balanced", synth.balanced_tree)):
        ft = gen(1_000_000, seed=3)
        t0 = time.perf_counter()
        T = SuchTree.from_flat(ft, device=local)
        build_s = time.perf_counter() - t0
        T.random_leaf_pairs_device(3, rank * n3, n3, pairs.data_ptr(), idx_bits=32, stream=sptr)
        sec = _timed(stream, lambda: T.distances_device(pairs.data_ptr(), n3, out.data_ptr(), idx_bits=32,
                                                        stream=sptr), steps=5)
        T.check_range(sptr)
        sec = max_over_ranks(sec)
        # size-independent properties at full size: symmetric pairs give d(a,a)=0 and the
        # result is finite and bounded by twice the largest root distance
        finite = bool(torch.isfinite(out).all().item())
        # gather roofline for an index of this size (random sectors over the same footprint)
        sps = C.c_double(0)
        gfrac = None
        if _lib.lib().st_bench_gather(local, int(T.index_info["index_bytes"]), 64, 3, C.byref(sps)) == 0 and sps.value > 0:
            gfrac = 2.0 * n3 / sec / sps.value
        res["cfg3_" + shape] = {
            "workload": "1M-leaf %s tree (depth %d), %d random leaf pairs per GPU per launch" % (shape, T.depth, n3),
            "pairs_per_s": world * n3 / sec, "ms_per_launch": sec * 1e3, "index_build_s": build_s,
            "index_bytes": int(T.index_info["index_bytes"]),
            "hbm_frac": 16.0 * n3 / sec / 1e9 / peak, "gather_frac": gfrac,
            "gather_peak_sectors_per_s": sps.value, "all_finite": finite,
        }
        del T
    del pairs, out
    torch.cuda.empty_cache()

    # ---- cfg5: 100k x 100k fp64 matrix, one row block per GPU, left on the device
    ft = synth.yule_tree(TREE_LEAVES, seed=TREE_SEED)
    T = SuchTree.from_flat(ft, device=local)
    n = TREE_LEAVES
    rb, re_ = shard.row_block(rank, max(world, 8), n)  # 12,500-row blocks as in the 8-GPU plan
    block = torch.empty((re_ - rb, n), dtype=torch.float64, device=dev)
    from suchtree_b200 import _lib

    def mat():
        _lib.check(_lib.lib().st_distance_matrix(T._handle, None, n, rb, re_, block.data_ptr(), 1, sptr))

    sec = max_over_ranks(_timed(stream, mat, steps=10))
    elems = (re_ - rb) * n
    # spot check against the pair kernel (same index): 4096 random elements of the block
    g = torch.Generator(device="cpu").manual_seed(11 + rank)
    ri = torch.randint(0, re_ - rb, (4096,), generator=g)
    ci = torch.randint(0, n, (4096,), generator=g)
    chk_pairs = torch.stack([2 * (ri + rb), 2 * ci], dim=1).to(torch.int32).to(dev)
    chk = torch.empty(4096, dtype=torch.float64, device=dev)
    T.distances_device(chk_pairs.data_ptr(), 4096, chk.data_ptr(), idx_bits=32, stream=sptr)
    torch.cuda.synchronize()
    same = bool(torch.equal(block[ri.to(dev), ci.to(dev)], chk))
    fill_sec = _timed(stream, lambda: block.fill_(1.0), steps=10)  # plain write of the same bytes
    res["cfg5_matrix"] = {
        "workload": "rows [%d,%d) of the 100k x 100k all-leaves fp64 matrix per GPU, device-resident" % (rb, re_),
        "elements_per_s": world * elems / sec, "ms_per_block": sec * 1e3, "bytes_written_per_gpu": 8 * elems,
        "hbm_frac": 8.0 * elems / sec / 1e9 / peak, "matches_pair_kernel_on_4096_samples": same,
        "plain_fill_gbs": 8.0 * elems / fill_sec / 1e9, "frac_of_plain_fill": fill_sec / sec,
    }
    del block
    torch.cuda.empty_cache()

    # ---- cfg4: sampled two-tree correlation, moments via NCCL
    fa, fb = synth.yule_tree(TREE_LEAVES, seed=4, names=True), synth.yule_tree(TREE_LEAVES, seed=5, names=True)
    TA, TB = SuchTree.from_flat(fa, device=local), SuchTree.from_flat(fb, device=local)
    rng = np.random.default_rng(6)
    la = 2 * rng.integers(0, TREE_LEAVES, TREE_LEAVES)
    lb = 2 * rng.integers(0, TREE_LEAVES, TREE_LEAVES)
    linklist = np.ascontiguousarray(np.stack([lb, la], axis=1).astype(np.int64))
    n4 = args.cfg4_samples
    from suchtree_b200 import _lib as L_

    import ctypes as C

    def sample():
        m = L_.Moments()
        L_.check(L_.lib().st_sample_moments(TA._handle, TB._handle, linklist.ctypes.data, linklist.shape[0], 7,
                                            rank * n4, n4, 0.0, 0.0, C.byref(m)))
        return m

    sample()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        m = sample()
    sec = max_over_ranks((time.perf_counter() - t0) / reps)
    m = shard.allreduce_moments(m, device=dev)
    res["cfg4_pearson_sampler"] = {
        "workload": "two 100k-leaf Yule trees (seeds 4,5), 100k random links (seed 6), %d Philox-sampled "
                    "link pairs per GPU per call, 5 moments all-reduced (NCCL)" % n4,
        "samples_per_s": world * n4 / sec, "ms_per_call": sec * 1e3, "pearson_r": moments_pearson(m),
        "timing": "host wall clock around the blocking C-ABI call (includes link upload + moment read-back)",
    }
    # ---- exhaustive link pairs, fused moments (linked_distances + pearson in one pass):
    #      44,904 links as in the reference's bigtrees example -> 1.008e9 link pairs,
    #      one eighth of the pair range per GPU
    L = 44_904
    ll2 = np.ascontiguousarray(linklist[:L])
    total = L * (L - 1) // 2
    pb, pe = shard.pair_range(rank, max(world, 8), total)

    def exhaustive():
        m = L_.Moments()
        L_.check(L_.lib().st_linked_moments(TA._handle, TB._handle, ll2.ctypes.data, L, pb, pe - pb, 0.0, 0.0,
                                            C.byref(m)))
        return m

    exhaustive()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        m2 = exhaustive()
    sec = max_over_ranks((time.perf_counter() - t0) / reps)
    res["linked_exhaustive_moments"] = {
        "workload": "44,904 links (1.008e9 link pairs): pairs [%d,%d) per GPU, both trees, moments fused" % (pb, pe),
        "link_pairs_per_s": world * (pe - pb) / sec, "ms_per_call": sec * 1e3,
        "timing": "host wall clock around the blocking C-ABI call",
    }
    del TA, TB

    # ---- the drop-in sampler API on the shape of the reference's published example
    #      (docs/examples/SuchLinkedTree_examples.md:76-149: 14 x 103,446 leaves, 44,904
    #      links; construction 4 min 57 s, sample_linked_distances 3.5e4 samples/s there)
    if rank == 0:
        import pandas as pd

        from suchtree_b200 import SuchLinkedTrees

        fa2 = synth.yule_tree(14, seed=8, names=True)
        fb2 = synth.yule_tree(103_446, seed=9, names=True)
        A2, B2 = SuchTree.from_flat(fa2, device=local), SuchTree.from_flat(fb2, device=local)
        rng2 = np.random.default_rng(10)
        mat = np.zeros((14, 103_446), dtype=np.int8)
        cols = rng2.choice(103_446, size=44_904, replace=False)
        mat[rng2.integers(0, 14, size=44_904), cols] = 1
        links = pd.DataFrame(mat, index=list(A2.leaves.keys()), columns=list(B2.leaves.keys()))
        t0 = time.perf_counter()
        SLT = SuchLinkedTrees(A2, B2, links)
        build_s = time.perf_counter() - t0
        SLT.sample_linked_distances(sigma=0.0, buckets=64, n=4096, maxcycles=1)  # warm
        t0 = time.perf_counter()
        cycles = 10
        SLT.sample_linked_distances(sigma=0.0, buckets=64, n=4096, maxcycles=cycles)  # never converges: 10 cycles
        dt = time.perf_counter() - t0
        res["sample_linked_distances_api"] = {
            "workload": "SuchLinkedTrees(14-leaf, 103,446-leaf trees, 44,904 links).sample_linked_distances("
                        "buckets=64, n=4096), %d cycles, the reference's exact xorshift64* stream" % cycles,
            "samples_per_s": cycles * 64 * 4096 / dt, "construction_s": build_s, "n_links": int(SLT.n_links),
        }
        del SLT, A2, B2

    # ---- N1: quartet topologies, 5e7 random leaf quartets per GPU, device resident
    T = SuchTree.from_flat(synth.yule_tree(TREE_LEAVES, seed=TREE_SEED), device=local)
    nq = args.quartets
    g = torch.Generator(device=dev).manual_seed(21 + rank)
    quartets = 2 * torch.randint(0, TREE_LEAVES, (nq, 4), generator=g, device=dev, dtype=torch.int64)
    topo = torch.empty_like(quartets)
    sec = max_over_ranks(_timed(stream, lambda: T.quartet_topologies_device(quartets.data_ptr(), nq, topo.data_ptr(),
                                                                            stream=sptr), steps=5))
    T.check_range(sptr)
    is_perm = bool(torch.equal(torch.sort(topo, dim=1).values, torch.sort(quartets, dim=1).values))
    res["n1_quartet_topologies"] = {
        "workload": "%d random leaf quartets per GPU per launch on the cfg2 tree (int64 x4 in, int64 x4 out)" % nq,
        "quartets_per_s": world * nq / sec, "ms_per_launch": sec * 1e3,
        "hbm_frac": 64.0 * nq / sec / 1e9 / peak, "rows_are_permutations": is_perm,
    }
    del T, quartets, topo

    # ---- the reference's per-clade scan (docs/examples/SuchLinkedTree_examples.md:286-310:
    #      for every internal node of TreeB { subset_b; linked_distances; pearson }, keeping
    #      clades with 10..2500 links; 6 h 39 min there) as one launch sequence.  Rank 0 only,
    #      no collective; on its own try so that a failure costs this entry alone.
    if rank == 0:
        try:
            from suchtree_b200 import SuchLinkedTrees

            fa3, fb3 = synth.yule_tree(TREE_LEAVES, seed=4, names=True), synth.yule_tree(TREE_LEAVES, seed=5, names=True)
            A3, B3 = SuchTree.from_flat(fa3, device=local), SuchTree.from_flat(fb3, device=local)
            SLT3 = SuchLinkedTrees.from_linklist(A3, B3, linklist)
            SLT3.clade_pearson(min_links=10, max_links=2500)  # warm: clade intervals, scratch pool
            t0 = time.perf_counter()
            reps3 = 3
            for _ in range(reps3):
                scan = SLT3.clade_pearson(min_links=10, max_links=2500)
            dt = (time.perf_counter() - t0) / reps3
            done = np.isfinite(scan["r"])
            res["clade_scan"] = {
                "workload": "clade_pearson(min_links=10, max_links=2500) over all %d internal nodes of TreeB: two "
                            "100k-leaf Yule trees, 100k random links (cfg4's), one GPU" % int(scan["node_ids"].shape[0]),
                "clades_computed": int(done.sum()), "link_pairs": int(scan["n_pairs"].sum()),
                "link_pairs_per_s": float(scan["n_pairs"].sum()) / dt, "s_per_scan": dt,
                "timing": "host wall clock around the Python call (link sort + work-item plan on the host, "
                          "3 kernels, moments read back)",
            }
            del SLT3, A3, B3
        except Exception as e:  # diagnostics only
            res["clade_scan"] = {"error": repr(e)[:200]}
    return res

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP, help="pairs per GPU per step")
    ap.add_argument("--e2e-pairs", type=int, default=PAIRS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the cfg3/cfg4/cfg5 side measurements")
    ap.add_argument("--cfg3-pairs", type=int, default=1_250_000_000, help="cfg3 pairs per GPU per launch")
    ap.add_argument("--cfg4-samples", type=int, default=125_000_000, help="cfg4 samples per GPU per call")
    ap.add_argument("--quartets", type=int, default=50_000_000, help="quartets per GPU per launch")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
