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
`e2e` = the same metric through the drop-in call with the REFERENCE'S signature:
r = SuchTree.distances_bulk(pairs), pairs an ordinary (pageable) numpy int64 (n,2)
array, r a fresh float64 array (H2D + kernel + D2H inside the timed region), with
`e2e.roofline` = a concurrent pinned H2D + D2H copy of the same byte counts measured
in the same run (all ranks at once for N>1).
Rank 0 at N=1 also times the unmodified reference (oracle/_ref, its own Cython
machine code) on a bounded sample of the same pairs -> `cpu_baseline`.

Reference arm (--impl reference): the unmodified reference's distances_bulk() on
the box's host cores (fork pool over contiguous pair blocks, the decomposition the
reference's docs recommend), every step the FULL 1e8-pair step of our arm.
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
_REF_OUT = None


def _ref_worker(span):
    # the pair array, the tree and the result array are inherited through fork (the result
    # array is shared anonymous memory): only the span travels, as cheap as it can be made
    # for the reference -- its own recipe returns the results through the pool's pipes
    _REF_OUT[span[0]:span[1]] = _REF_TREE.distances_bulk(_REF_PAIRS[span[0]:span[1]])
    return span[1] - span[0]


class ReferenceCPU:
    """The unmodified reference extension (oracle/_ref) on the bench tree, or the C
    port (oracle/st_oracle.c, literal fp32 arithmetic) if the extension is absent."""

    def __init__(self, flat_tree):
        global _REF_TREE
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        self.kind = "port"
        self.pool = None
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
        pool, out = self.open_pool(pairs, cores)
        try:
            return pairs.shape[0] / self.run_pool(pairs.shape[0], cores), out.copy()
        finally:
            self.close_pool()

    def open_pool(self, pairs, cores):
        """fork Pool(cores) that sees `pairs` and a shared result array (threads do not
        scale: the reference holds the GIL in distances_bulk)."""
        import mmap
        import multiprocessing as mp

        global _REF_PAIRS, _REF_OUT
        _REF_PAIRS = pairs
        self._shm = mmap.mmap(-1, max(8, 8 * pairs.shape[0]))  # MAP_SHARED | MAP_ANONYMOUS: survives fork
        _REF_OUT = np.frombuffer(self._shm, dtype=np.float64, count=pairs.shape[0])
        self.pool = mp.get_context("fork").Pool(cores)
        self.pool.map(_ref_worker, [(0, min(1000, pairs.shape[0]))] * cores)  # warm the workers
        return self.pool, _REF_OUT

    def run_pool(self, n, cores):
        """seconds for one pass over the n pairs given to open_pool()"""
        edges = np.linspace(0, n, cores * 4 + 1).astype(np.int64)
        spans = list(zip(edges[:-1].tolist(), edges[1:].tolist()))
        t0 = time.perf_counter()
        done = sum(self.pool.map(_ref_worker, spans))
        dt = time.perf_counter() - t0
        assert done == n
        return dt

    def close_pool(self):
        global _REF_PAIRS, _REF_OUT
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
        _REF_PAIRS = _REF_OUT = None


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
    probe = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, 200_000)
    rate1, _ = ref.run(probe, 1)
    # the same step as our arm: rank 0's 1e8 pairs of the Philox stream (all of them, every step)
    per_step = args.pairs - args.pairs % 2
    pairs = sample_pairs_host(TREE_LEAVES, PAIR_SEED, 0, per_step)
    ref.open_pool(pairs, cores)
    try:
        for _ in range(min(max(args.warmup, 0), 1)):  # one warm pass is plenty for a CPU loop
            ref.run_pool(per_step, cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.run_pool(per_step, cores)
        dt = time.perf_counter() - t0
    finally:
        ref.close_pool()
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
        "config": {"workload": WORKLOAD, "tree_leaves": TREE_LEAVES, "tree_nodes": int(ft.size),
                   "pairs_per_step_per_gpu": per_step, "pair_dtype": "int64x2", "result_dtype": "f64",
                   "note": "the same step as our arm: all 1e8 pairs of rank 0's Philox stream, every step (the "
                           "CPU arm does not grow with --gpus); pool start-up outside the rate; results written "
                           "to shared memory by the workers"},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": ref.kind, "cpu_model": cpu_model(),
            "sample": "%d pairs/step x %d steps of the cfg2 pair stream (the whole step), fork Pool(%d) over "
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


def ncu_traffic(build_id, n_pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the
    committed `ncu --set full` capture (profiles/traffic.json) -- only if that capture was
    taken from THIS pair kernel (same st_pairs_kernel_id: hash of its sources and compiler
    flags) and launch size; otherwise null, with the reason."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
    except Exception:
        return None, "no profiles/traffic.json"
    if t.get("pairs_kernel_id") != build_id:
        return None, "stale: captured from kernel %s, this is %s (%s)" % (t.get("pairs_kernel_id"), build_id, t.get("capture"))
    if int(t.get("pairs_per_launch", 0)) != int(n_pairs):
        return None, "captured at %s pairs per launch" % t.get("pairs_per_launch")
    return t.get("k_pairs_dram_bytes_per_launch"), "ncu capture %s (pair kernel %s)" % (t.get("capture"), build_id)


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

    # ---- e2e: the drop-in call with the reference's own signature (MuchTree.pyx:872-909):
    #      an ordinary pageable numpy int64 (n,2) array in, a FRESH float64 array out
    import ctypes as C

    n_e2e = min(n_pairs, args.e2e_pairs)
    np_pairs = pairs[:n_e2e].cpu().numpy().astype(np.int64)  # pageable host memory, C-contiguous
    assert np_pairs.flags.owndata and not _lib.lib().st_host_is_pinned(np_pairs.ctypes.data)
    t0 = time.perf_counter()
    r = T.distances_bulk(np_pairs)  # cold call: result pool and (maybe) staging are allocated here
    e2e_first_s = time.perf_counter() - t0
    e2e_ok = bool(torch.equal(torch.from_numpy(r).to(dev), out[:n_e2e]))
    for _ in range(max(args.warmup, 3) - 1):
        r = T.distances_bulk(np_pairs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = T.distances_bulk(np_pairs)
    dt = time.perf_counter() - t0
    e2e_ok = e2e_ok and bool(torch.equal(torch.from_numpy(r).to(dev), out[:n_e2e]))
    in_registered = bool(_lib.lib().st_host_is_pinned(np_pairs.ctypes.data))

    def all_max(x):
        t_ = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    e2e_value = world * n_e2e * args.steps / all_max(dt)

    # variants, a few calls each (not the headline): (a) inputs never page-locked -- what a caller
    # that hands in a NEW array every call gets; (b) round 1's extension: pinned in, pinned out=
    variants = {}
    reps = max(3, min(args.steps, 5))
    os.environ["SUCHTREE_B200_REGISTER"] = "0"
    np_pairs2 = np_pairs.copy()
    r = T.distances_bulk(np_pairs2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = T.distances_bulk(np_pairs2)
    variants["pageable_in_never_registered"] = world * n_e2e * reps / all_max(time.perf_counter() - t0)
    del os.environ["SUCHTREE_B200_REGISTER"]
    del np_pairs2
    h_pairs = torch.empty((n_e2e, 2), dtype=torch.int64).pin_memory()
    h_pairs.copy_(torch.from_numpy(np_pairs))
    h_out = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    T.distances_bulk(h_pairs.numpy(), out=h_out.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        T.distances_bulk(h_pairs.numpy(), out=h_out.numpy())
    variants["pinned_in_pinned_out_kwarg"] = world * n_e2e * reps / all_max(time.perf_counter() - t0)
    del h_pairs, h_out, r

    # the roofline of that call: what the host interface carries when NOTHING else happens --
    # a pinned H2D copy of 16 B/pair and a concurrent pinned D2H copy of 8 B/pair, same chunking,
    # all ranks at the same time (the ranks of one box share the host's memory system)
    sec = C.c_double(0)
    barrier()
    rc_copy = _lib.lib().st_bench_copy(local, 16 * n_e2e, 8 * n_e2e, 64 << 20, 5, C.byref(sec))
    copy_s = all_max(sec.value if rc_copy == 0 else float("nan"))
    copy_pairs_per_s = world * n_e2e / copy_s
    e2e_roofline = {
        "bound": "host interface (PCIe Gen5 x16 per GPU; the host memory system when several ranks share it)",
        "what": "cudaMemcpyAsync pinned H2D of 16 B/pair || pinned D2H of 8 B/pair, 64 MiB chunks, two streams, "
                "%d rank(s) concurrently, measured in this run by st_bench_copy" % world,
        "achieved": 24.0 * e2e_value / 1e9, "peak": 24.0 * copy_pairs_per_s / 1e9, "unit": "GB/s",
        "frac": e2e_value / copy_pairs_per_s,
        "peak_pairs_per_s": copy_pairs_per_s, "algorithmic_bytes_per_pair": 24,
        "note": "frac can exceed 1: the pipeline bit-packs part of every chunk's ids on the host (2 x ceil(log2 "
                "n_nodes) bits per pair; 70 % of a chunk on one GPU, 20 % with two local ranks, 0 % beyond), so "
                "fewer bytes cross PCIe than the plain copy moves; `packed_bound` is the same measurement with "
                "every id packed -- what PCIe alone would allow if packing cost the host nothing",
    }
    # the other end of the bracket: H2D of fully bit-packed ids || D2H of the results
    _pf, _ib = C.c_double(0), C.c_int(0)
    _lib.check(_lib.lib().st_host_route_info(T._handle, C.byref(_pf), C.byref(_ib)))
    route_frac, route_bits = (float(_pf.value) if in_registered else 1.0), int(_ib.value)
    id_bits = route_bits
    sec2 = C.c_double(0)
    barrier()
    rc_copy2 = _lib.lib().st_bench_copy(local, (2 * id_bits * n_e2e + 7) // 8, 8 * n_e2e, 64 << 20, 5, C.byref(sec2))
    copy2_s = all_max(sec2.value if rc_copy2 == 0 else float("nan"))
    e2e_roofline["packed_bound"] = {
        "what": "cudaMemcpyAsync pinned H2D of %.2f B/pair (2 x %d bits) || pinned D2H of 8 B/pair, same chunking and "
                "ranks" % (2 * id_bits / 8.0, id_bits),
        "peak_pairs_per_s": world * n_e2e / copy2_s, "frac": e2e_value / (world * n_e2e / copy2_s),
    }

    # ---- rooflines (rank 0's kernel; all ranks run the same launch)
    peak, peak_src = measured_peak()
    per_launch_s = ms * 1e-3 / args.steps
    achieved = 16.0 * n_pairs / per_launch_s / 1e9
    traffic, traffic_src = ncu_traffic(_lib.lib().st_pairs_kernel_id().decode(), n_pairs)
    gather = None
    if rank == 0:
        sps = C.c_double(0)
        rc = _lib.bench_lib().st_bench_gather(local, int(T.index_info["index_bytes"]), 64, 5, C.byref(sps))
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
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": "k_pairs<int32, 2 pairs/thread, compact layout, 512 threads x 2 CTAs/SM>",
            "algorithmic_bytes_per_pair": 16,
            "note": "the kernel is bound by the per-SM L1TEX rate of random sector gathers (2 per pair), see gather_roofline",
        },
        "gather_roofline": gather,
        "cpu_baseline": cpu_baseline,
        "e2e": {
            "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n_e2e, "d2h_bytes_per_step": 8 * n_e2e,
            "h2d_bytes_over_pcie_per_step": int(n_e2e * (route_frac * 2 * route_bits / 8.0 + (1.0 - route_frac) * 16)),
            "pack_fraction": route_frac, "id_bits": route_bits,
            "pairs_per_step_per_gpu": n_e2e, "steps": args.steps, "matches_device_path": e2e_ok,
            "call": "r = SuchTree.distances_bulk(pairs): pairs = ordinary numpy int64 (n,2) array (the same "
                    "array every step), r = fresh numpy float64 (n,) -- the reference's signature, "
                    "MuchTree.pyx:872-909",
            "input_registered_in_place": in_registered,
            "first_call_s": e2e_first_s,
            "roofline": e2e_roofline,
            "variants_pairs_per_s": variants,
        },
        "gpu_launches": args.steps,
        "clocks": clocks,
        "parity_vs_reference_sample": parity,
    }
    if not args.no_other_workloads:
        del pairs, out, np_pairs
        _lib.lib().st_host_trim(0)
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


def cfg5_parity_vs_oracle(ft, T, block, n, n_shards, mat, dev):
    """The cfg5 matrix against the oracle at full size (1e6 sampled elements +
    row checksums).  Every row shard of the 8-GPU plan is computed into `block` in turn."""
    import torch

    from suchtree_b200 import shard

    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle as O

    ot = O.OracleTree(ft.parent, ft.distance)
    rng = np.random.default_rng(55)
    TR, TC = 64, 512  # the writer's tile (st_matrix.cu)
    n_checked = 0
    bad = 0
    rows_checked = 0
    row_bad = 0
    for sh in range(n_shards):
        r0, r1 = shard.row_block(sh, n_shards, n)
        if r1 <= r0:
            continue
        mat(r0, r1)
        torch.cuda.synchronize()
        rows = r1 - r0
        k = 1_000_000 // n_shards
        ri = [rng.integers(0, rows, k - 6000)]
        ci = [rng.integers(0, n, k - 6000)]
        # elements of the tiles ON the diagonal and just off it on both sides
        d = rng.integers(0, rows, 3000)
        off = rng.integers(-TC, TC + 1, 3000)
        ri.append(d)
        ci.append(np.clip(r0 + d + off, 0, n - 1))
        # the last (partial) tile column and the shard's first / last tile rows
        e = rng.integers(0, rows, 1500)
        ri.append(e)
        ci.append(rng.integers((n - 1) // TC * TC, n, 1500))
        f = np.concatenate([rng.integers(0, min(TR, rows), 750), rng.integers(max(rows - TR, 0), rows, 750)])
        ri.append(f)
        ci.append(rng.integers(0, n, 1500))
        ri, ci = np.concatenate(ri), np.concatenate(ci)
        got = block[torch.from_numpy(ri).to(dev), torch.from_numpy(ci).to(dev)].cpu().numpy()
        want = ot.distances_f64_climb(np.stack([2 * (r0 + ri), 2 * ci], axis=1).astype(np.int64))
        bad += int(np.count_nonzero(got != want))
        n_checked += int(ri.shape[0])
        # whole rows: every element and the row sum (first, last and two random rows of the shard)
        for rr in sorted({0, rows - 1, int(rng.integers(0, rows)), int(rng.integers(0, rows))}):
            row = block[rr].cpu().numpy()
            p = np.stack([np.full(n, 2 * (r0 + rr)), 2 * np.arange(n)], axis=1).astype(np.int64)
            w = ot.distances_f64_climb(p)
            rows_checked += 1
            if not (np.array_equal(row, w) and float(row.sum()) == float(w.sum())):
                row_bad += 1
    return {"oracle": "O2 (oracle/st_oracle.c: fp64 path sums over the fp32-quantised edges, climbing MRCA)",
            "shards": n_shards, "sampled_elements": n_checked, "sampled_mismatches": bad,
            "whole_rows": rows_checked, "row_mismatches": row_bad, "bit_exact": bool(bad == 0 and row_bad == 0)}


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
        for _ in range(3):  # (two result blocks alternate while `got` is alive: both exist after this)
            got = G.distances_bulk(p1)
        t0 = time.perf_counter()
        for _ in range(10):
            got = G.distances_bulk(p1)
        host_s = (time.perf_counter() - t0) / 10
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
        if _lib.bench_lib().st_bench_gather(local, int(T.index_info["index_bytes"]), 64, 3, C.byref(sps)) == 0 and sps.value > 0:
            gfrac = 2.0 * n3 / sec / sps.value
        res["cfg3_" + shape] = {
            "workload": "1M-leaf %s tree (depth %d), %d random leaf pairs per GPU per launch" % (shape, T.depth, n3),
            "pairs_per_s": world * n3 / sec, "ms_per_launch": sec * 1e3, "index_build_s": build_s,
            "index_bytes": int(T.index_info["index_bytes"]),
            "hbm_frac": 16.0 * n3 / sec / 1e9 / peak, "gather_frac": gfrac,
            "gather_peak_sectors_per_s": sps.value, "all_finite": finite,
            "paired_record_kernel": bool(T.index_info["paired_records"]),
            "probe_neighbour_is_mrca": float(T.index_info["probe_neighbour_hit"]),
        }
        del T
    del pairs, out
    torch.cuda.empty_cache()

    # ---- cfg5: 100k x 100k fp64 matrix, one row block per GPU, left on the device
    ft = synth.yule_tree(TREE_LEAVES, seed=TREE_SEED)
    T = SuchTree.from_flat(ft, device=local)
    n = TREE_LEAVES
    n_shards = max(world, 8)  # 12,500-row blocks as in the 8-GPU plan
    rb, re_ = shard.row_block(rank, n_shards, n)
    block = torch.empty((re_ - rb, n), dtype=torch.float64, device=dev)

    def mat(r0=rb, r1=re_):
        _lib.check(_lib.lib().st_distance_matrix(T._handle, None, n, r0, r1, block.data_ptr(), 1, sptr))

    sec = max_over_ranks(_timed(stream, mat, steps=10))
    elems = (re_ - rb) * n
    fill_sec = _timed(stream, lambda: block.fill_(1.0), steps=10)  # plain write of the same bytes
    res["cfg5_matrix"] = {
        "workload": "rows [%d,%d) of the 100k x 100k all-leaves fp64 matrix per GPU, device-resident" % (rb, re_),
        "elements_per_s": world * elems / sec, "ms_per_block": sec * 1e3, "bytes_written_per_gpu": 8 * elems,
        "hbm_frac": 8.0 * elems / sec / 1e9 / peak,
        "plain_fill_gbs": 8.0 * elems / fill_sec / 1e9, "frac_of_plain_fill": fill_sec / sec,
    }
    # full-size parity against the oracle (the checker; rank 0, N=1, beside the cpu_baseline leg):
    # every one of the 8 row shards is computed in turn; 1e6 sampled elements in all -- drawn
    # uniformly, plus the diagonal tiles, both triangle sides next to the diagonal, the last
    # (partial) tile column / row -- and whole-row checksums, against O2 (fp64 path sums over the
    # fp32-quantised edges, the climbing-MRCA restatement), bit for bit.
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res["cfg5_matrix"]["parity_vs_oracle"] = cfg5_parity_vs_oracle(ft, T, block, n, n_shards, mat, dev)
    del block
    torch.cuda.empty_cache()

    # ---- the drop-in call itself, to the host: pairwise_distances() over 20,000 leaves (3.2 GB),
    #      about the most the reference's own implementation (n(n-1)/2 Python tuples) can attempt
    if rank == 0:
        nodes = list(range(0, 40_000, 2))
        t0 = time.perf_counter()
        D = T.pairwise_distances(nodes)  # cold: the pinned result block is allocated here
        cold = time.perf_counter() - t0
        del D
        t0 = time.perf_counter()
        D = T.pairwise_distances(nodes)
        warm = time.perf_counter() - t0
        sym = bool(np.array_equal(D[:64, :], D[:, :64].T)) and bool(np.all(np.diagonal(D) == 0.0))
        chk = T.distances_bulk(np.array([[nodes[17], nodes[19_999]], [nodes[12_345], nodes[6_789]]], dtype=np.int64))
        res["pairwise_distances_host_20k"] = {
            "workload": "SuchTree.pairwise_distances(20,000 leaf ids) -> fresh (20000,20000) float64 numpy array "
                        "(3.2 GB) in host memory",
            "s_first_call": cold, "s_per_call": warm, "elements_per_s": 4e8 / warm, "d2h_gbs": 3.2 / warm,
            "symmetric_zero_diagonal": sym,
            "matches_distances_bulk": bool(D[17, 19_999] == chk[0] and D[12_345, 6_789] == chk[1]),
        }
        del D
        _lib.lib().st_host_trim(0)

    # ---- cfg4: sampled two-tree correlation; the six moment sums are all-reduced by NCCL on the
    #      kernel's stream INSIDE the timed call (st_links_sample_moments(..., nccl_comm))
    fa, fb = synth.yule_tree(TREE_LEAVES, seed=4, names=True), synth.yule_tree(TREE_LEAVES, seed=5, names=True)
    TA, TB = SuchTree.from_flat(fa, device=local), SuchTree.from_flat(fb, device=local)
    rng = np.random.default_rng(6)
    la = 2 * rng.integers(0, TREE_LEAVES, TREE_LEAVES)
    lb = 2 * rng.integers(0, TREE_LEAVES, TREE_LEAVES)
    linklist = np.ascontiguousarray(np.stack([lb, la], axis=1).astype(np.int64))
    n4 = args.cfg4_samples
    L_ = _lib
    links = C.c_void_p()
    L_.check(L_.lib().st_links_create(TA._handle, TB._handle, linklist.ctypes.data, linklist.shape[0], C.byref(links)))
    comm = shard.MomentComm(local) if world > 1 else None
    comm_h = comm.handle if comm is not None else None

    def sample(first=rank * n4, count=n4, c=comm_h):
        m = L_.Moments()
        L_.check(L_.lib().st_links_sample_moments(links, 7, first, count, 0.0, 0.0, c, C.byref(m)))
        return m

    for _ in range(3):
        sample()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        m = sample()
    sec = max_over_ranks((time.perf_counter() - t0) / reps)
    r_all = moments_pearson(m)
    entry = {
        "workload": "two 100k-leaf Yule trees (seeds 4,5), 100k random links (seed 6), %d Philox-sampled link "
                    "pairs per GPU per call; %s" % (
                        n4, "moments all-reduced by ncclAllReduce(sum, fp64, 6) on the kernel's stream inside "
                            "every timed call" if world > 1 else "one GPU: no collective"),
        "samples_per_s": world * n4 / sec, "ms_per_call": sec * 1e3, "pearson_r": r_all, "samples_in_r": m.n,
        "timing": "host wall clock around the blocking C-ABI call (kernel + fold + all-reduce + 48-byte read-back); "
                  "the link list is device-resident (st_links handle)",
    }
    if world > 1:
        # the all-reduced r must equal ONE GPU's r over the union of all ranks' sample ranges
        m1 = sample(0, world * n4, None)
        r_one = moments_pearson(m1)
        entry["allreduce_check"] = {
            "r_allreduced": r_all, "r_one_gpu_union_range": r_one, "abs_diff": abs(r_all - r_one),
            "ok_1e-12": bool(abs(r_all - r_one) <= 1e-12 and m.n == m1.n),
        }
        # what the collective costs: the same call without it
        t0 = time.perf_counter()
        for _ in range(reps):
            sample(c=None)
        entry["ms_per_call_without_allreduce"] = max_over_ranks((time.perf_counter() - t0) / reps) * 1e3
    res["cfg4_pearson_sampler"] = entry
    # ---- exhaustive link pairs, fused moments (linked_distances + pearson in one pass):
    #      44,904 links as in the reference's bigtrees example -> 1.008e9 link pairs,
    #      one eighth of the pair range per GPU
    L = 44_904
    ll2 = np.ascontiguousarray(linklist[:L])
    links2 = C.c_void_p()
    L_.check(L_.lib().st_links_create(TA._handle, TB._handle, ll2.ctypes.data, L, C.byref(links2)))
    total = L * (L - 1) // 2
    pb, pe = shard.pair_range(rank, max(world, 8), total)

    def exhaustive():
        m = L_.Moments()
        L_.check(L_.lib().st_links_linked_moments(links2, pb, pe - pb, 0.0, 0.0, comm_h, C.byref(m)))
        return m

    exhaustive()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        m2 = exhaustive()
    sec = max_over_ranks((time.perf_counter() - t0) / reps)
    res["linked_exhaustive_moments"] = {
        "workload": "44,904 links (1.008e9 link pairs): pairs [%d,%d) per GPU, both trees, moments fused%s" % (
            pb, pe, ", all-reduced inside the call" if world > 1 else ""),
        "link_pairs_per_s": world * (pe - pb) / sec, "ms_per_call": sec * 1e3,
        "timing": "host wall clock around the blocking C-ABI call",
    }
    L_.lib().st_links_destroy(links)
    L_.lib().st_links_destroy(links2)
    if comm is not None:
        comm.close()
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

    # ---- the reference's own published benchmark (docs/benchmarks.md:31-76): one million random
    #      leaf-NAME pairs through distances_by_name() on the ML and the NJ tree of the same 54,327
    #      taxa (108,653 nodes each), then Pearson r of the two distance vectors.  Published: 10.1 s
    #      for the two calls on an i7-3770S (1.98e5 pairs/s), r = 0.969.
    if rank == 0:
        try:
            import gzip
            import random

            def _tree(name):
                with gzip.open(os.path.join(REPO, "tests", "golden", "data", name), "rt") as f:
                    return SuchTree(f.read().strip(), device=local)

            t0 = time.perf_counter()
            T1, T2 = _tree("ml.tree.gz"), _tree("nj.tree.gz")
            load_s = time.perf_counter() - t0
            random.seed(12)
            v = list(T1.leaves.keys())
            npairs = 1_000_000
            name_pairs = [(random.choice(v), random.choice(v)) for _ in range(npairs)]
            T1.distances_by_name(name_pairs[:1000])
            t0 = time.perf_counter()
            D1 = T1.distances_by_name(name_pairs)
            D2 = T2.distances_by_name(name_pairs)
            dt = time.perf_counter() - t0
            from suchtree_b200 import pearson as st_pearson

            r = st_pearson(np.asarray(D1), np.asarray(D2), device=local)
            res["published_benchmark_by_name"] = {
                "workload": "docs/benchmarks.md: 1e6 random leaf-name pairs, distances_by_name() on ml.tree and "
                            "nj.tree (54,327 leaves, 108,653 nodes each), names -> ids included",
                "s_both_calls": dt, "pairs_per_s": 2 * npairs / dt, "load_two_trees_s": load_s,
                "published_s": 10.1, "published_pairs_per_s": 1.98e5, "published_hardware": "Intel i7-3770S, 1 thread",
                "pearson_r": r, "published_pearson_r": 0.969,
                "nodes": [T1.size, T2.size], "leaves": [T1.num_leaves, T2.num_leaves],
                "layouts": [int(T1.index_info["layout"]), int(T2.index_info["layout"])],
            }
            del T1, T2, D1, D2, name_pairs
        except Exception as e:  # diagnostics only
            res["published_benchmark_by_name"] = {"error": repr(e)[:200]}

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
                "timing": "host wall clock around the Python call (plan, shift, moments and fold kernels on the "
                          "device; sorted links and prefix counts kept in the st_links handle; moments read back)",
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
