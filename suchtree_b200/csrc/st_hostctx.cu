// Per-device host context: lanes (streams + staging + status word per in-flight host
// call), the pinned result pool behind the Python shim's fresh result arrays, the
// registration helpers for large caller inputs, and the copy roofline of the host path.
#include "st_hostctx.cuh"
#include "st_hostpool.cuh"

#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

const int ST_MAX_DEVICES = 64;
const int ST_MAX_LANES = 4;  // concurrent host calls per device; further callers wait

struct DeviceCtx {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<HostLane *> lanes;
    bool prewarm_started = false, prewarming = false;
    std::thread prewarm_thread;
};

DeviceCtx g_ctx[ST_MAX_DEVICES];

// joins the prewarm threads before the CUDA runtime is torn down at exit
struct Joiner {
    ~Joiner() {
        for (auto &c : g_ctx)
            if (c.prewarm_thread.joinable()) c.prewarm_thread.join();
    }
} g_joiner;

int lane_create(int device, HostLane **out) {
    HostLane *l = new HostLane();
    l->device = device;
    for (int i = 0; i < ST_LANE_SLOTS; ++i) {
        ST_CUDA(cudaStreamCreateWithFlags(&l->streams[i], cudaStreamNonBlocking));
        ST_CUDA(cudaEventCreateWithFlags(&l->ev[i], cudaEventDisableTiming));
    }
    ST_CUDA(cudaMalloc(reinterpret_cast<void **>(&l->d_status), sizeof(RangeStatus)));
    ST_CUDA(cudaMemset(l->d_status, 0, sizeof(RangeStatus)));
    ST_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&l->h_status), sizeof(RangeStatus), cudaHostAllocPortable));
    ST_CUDA(cudaHostAlloc(&l->h_small_in, size_t(ST_MEDIUM_CALL) * 8, cudaHostAllocPortable));
    ST_CUDA(cudaHostAlloc(&l->h_small_out, size_t(ST_MEDIUM_CALL) * 8, cudaHostAllocPortable));
    ST_CUDA(cudaMalloc(reinterpret_cast<void **>(&l->d_scratch), 64 * sizeof(double)));
    ST_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&l->h_scratch), 64 * sizeof(double), cudaHostAllocPortable));
    *out = l;
    return ST_OK;
}

}  // namespace

HostLane *st_lane_acquire(int device) {
    if (device < 0 || device >= ST_MAX_DEVICES) {
        st_set_error("device %d out of range", device);
        return nullptr;
    }
    DeviceCtx &c = g_ctx[device];
    std::unique_lock<std::mutex> lk(c.mu);
    for (;;) {
        if (!c.prewarming) {
            for (HostLane *l : c.lanes)
                if (!l->busy) {
                    l->busy = true;
                    return l;
                }
            if (int(c.lanes.size()) < ST_MAX_LANES) {
                HostLane *l = nullptr;
                if (lane_create(device, &l) != ST_OK) return nullptr;  // (a half-built lane is leaked: the device is unusable anyway)
                l->busy = true;
                c.lanes.push_back(l);
                return l;
            }
        }
        c.cv.wait(lk);
    }
}

void st_lane_release(HostLane *lane) {
    DeviceCtx &c = g_ctx[lane->device];
    {
        std::lock_guard<std::mutex> lk(c.mu);
        lane->busy = false;
    }
    c.cv.notify_all();
}

int st_lane_ensure_stage(HostLane *l, int64_t n, bool need_h_in, bool need_h_out) {
    int64_t want = 4096;
    while (want < n && want < ST_STAGE_PAIRS_MAX) want <<= 1;
    if (want > l->stage_pairs) {
        for (int i = 0; i < ST_LANE_SLOTS; ++i) cudaStreamSynchronize(l->streams[i]);
        for (int i = 0; i < ST_LANE_SLOTS; ++i) {
            cudaFree(l->d_in[i]);
            cudaFree(l->d_out[i]);
            cudaFree(l->d_out2[i]);
            if (l->h_in[i]) st_host_free(l->h_in[i]);
            if (l->h_out[i]) st_host_free(l->h_out[i]);
            l->d_in[i] = l->d_out[i] = l->d_out2[i] = l->h_in[i] = l->h_out[i] = nullptr;
        }
        l->stage_pairs = 0;
        for (int i = 0; i < ST_LANE_SLOTS; ++i) {
            ST_CUDA(cudaMalloc(&l->d_in[i], size_t(want) * 16));
            ST_CUDA(cudaMalloc(&l->d_out[i], size_t(want) * 8));
            ST_CUDA(cudaMalloc(&l->d_out2[i], size_t(want) * 4));
        }
        l->stage_pairs = want;
    }
    for (int i = 0; i < ST_LANE_SLOTS; ++i) {
        int rc = ST_OK;
        if (need_h_in && !l->h_in[i] && (rc = st_pinned_alloc(size_t(l->stage_pairs) * 16, &l->h_in[i])) != ST_OK) return rc;
        if (need_h_out && !l->h_out[i] && (rc = st_pinned_alloc(size_t(l->stage_pairs) * 8, &l->h_out[i])) != ST_OK) return rc;
    }
    return ST_OK;
}

int st_lane_read_status(HostLane *l, cudaStream_t stream, unsigned long long *max_bad, long long *min_bad) {
    ST_CUDA(cudaMemcpyAsync(l->h_status, l->d_status, sizeof(RangeStatus), cudaMemcpyDeviceToHost, stream));
    ST_CUDA(cudaStreamSynchronize(stream));
    *max_bad = l->h_status->max_bad;
    *min_bad = l->h_status->min_bad;
    if (*max_bad != 0 || *min_bad != 0) {
        ST_CUDA(cudaMemsetAsync(l->d_status, 0, sizeof(RangeStatus), stream));
        ST_CUDA(cudaStreamSynchronize(stream));
    }
    return ST_OK;
}

void st_hostctx_prewarm(int device) {
    if (device < 0 || device >= ST_MAX_DEVICES) return;
    if (const char *e = getenv("SUCHTREE_B200_PREWARM"))
        if (e[0] == '0') return;
    DeviceCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.prewarm_started) return;
    c.prewarm_started = true;
    c.prewarming = true;
    c.prewarm_thread = std::thread([device, &c] {
        HostLane *l = nullptr;
        if (cudaSetDevice(device) == cudaSuccess && lane_create(device, &l) == ST_OK) {
            // full-size device staging and pinned input staging; pinned result staging is
            // only needed by C callers that hand in pageable result buffers (lazy)
            if (st_lane_ensure_stage(l, ST_STAGE_PAIRS_MAX, true, false) != ST_OK) cudaGetLastError();
        }
        {
            std::lock_guard<std::mutex> lk2(c.mu);
            if (l) c.lanes.push_back(l);
            c.prewarming = false;
        }
        c.cv.notify_all();
    });
}

bool st_is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int st_raise_smem_impl(const void *kern, int device, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> configured;
    std::lock_guard<std::mutex> l(mu);
    int &c = configured[std::make_pair(kern, device)];
    if (c < bytes) {
        ST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        c = bytes;
    }
    return ST_OK;
}

// ------------------------------------------------------ pinned result pool --
// distances_bulk() returns a FRESH array (MuchTree.pyx:907).  A fresh pageable array of
// 8e8 bytes costs page faults plus a staged copy out of pinned memory; a block of this
// pool is page-locked already, so the D2H copies land in it directly.  Freed blocks are
// cached (same-shape calls in a loop recycle two blocks) up to a byte cap.
namespace {
// One page-locked block.  Large blocks are anonymous mappings on transparent huge pages,
// first-touched by the host pool and then page-locked in place (cudaHostRegister): 0.04-0.09 s
// for 800 MB against 0.31-0.41 s for cudaHostAlloc on the same box (scripts/pin_exp.cu,
// profiles/r02_summary.md) -- the cost a first large call pays.  Small blocks, and boxes where
// the mapping or the registration fails, use cudaHostAlloc.
struct PinnedBlk {
    size_t cls = 0;        // size class (bytes handed out)
    void *map = nullptr;   // mmap base (NULL: the block came from cudaHostAlloc)
    size_t map_bytes = 0;
};
struct PinnedPool {
    std::mutex mu;
    std::unordered_map<void *, PinnedBlk> live;           // handed out
    std::multimap<size_t, std::pair<void *, PinnedBlk>> free_blocks;  // cached: class bytes -> block
    size_t cached_bytes = 0;
    size_t cap_bytes = size_t(4) << 30;
    bool cap_read = false;
} g_pool;

size_t size_class(size_t bytes) {
    const size_t min_step = size_t(2) << 20;
    size_t p = min_step;
    while ((p << 1) <= bytes) p <<= 1;
    const size_t step = std::max(min_step, p >> 3);
    return (bytes + step - 1) / step * step;
}

const size_t ST_HUGE = size_t(2) << 20;
const size_t ST_MAP_MIN = size_t(4) << 20;  // two huge pages; below this cudaHostAlloc is as quick

bool map_mode_enabled() {
    static const bool on = [] {
        const char *e = getenv("SUCHTREE_B200_PINNED_MMAP");
        return !(e && e[0] == '0');
    }();
    return on;
}

// page-locked block of `cls` bytes (a multiple of 2 MiB); false: the caller falls back
bool map_pinned(size_t cls, void **out, PinnedBlk *blk) {
    const size_t map_bytes = cls + ST_HUGE;
    void *base = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return false;
    char *q = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(base) + ST_HUGE - 1) & ~uintptr_t(ST_HUGE - 1));
    madvise(q, cls, MADV_HUGEPAGE);  // advisory: 4 KiB pages still work, only slower to lock
    const int parts = int(std::min<size_t>(size_t(st_host_threads()), cls / (size_t(8) << 20) + 1));
    st_parallel_for(parts, [&](int p, int np) {  // first touch in parallel (zero pages are faulted in here)
        const size_t b = (cls / 4096) * size_t(p) / size_t(np) * 4096, e = (cls / 4096) * size_t(p + 1) / size_t(np) * 4096;
        for (size_t o = b; o < e; o += 4096) q[o] = 0;
    });
    if (cudaHostRegister(q, cls, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(base, map_bytes);
        return false;
    }
    blk->cls = cls;
    blk->map = base;
    blk->map_bytes = map_bytes;
    *out = q;
    return true;
}

void release_pinned(void *p, const PinnedBlk &b) {
    if (b.map) {
        if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();  // (at interpreter exit the runtime may be gone)
        munmap(b.map, b.map_bytes);
    } else if (cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
    }
}
}  // namespace

int st_pinned_alloc(size_t bytes, void **out) {
    *out = nullptr;
    const size_t cls = size_class(std::max<size_t>(bytes, 1));
    {
        std::lock_guard<std::mutex> l(g_pool.mu);
        auto it = g_pool.free_blocks.find(cls);
        if (it != g_pool.free_blocks.end()) {
            *out = it->second.first;
            g_pool.live[*out] = it->second.second;
            g_pool.free_blocks.erase(it);
            g_pool.cached_bytes -= cls;
            return ST_OK;
        }
    }
    void *p = nullptr;
    PinnedBlk blk;
    blk.cls = cls;
    if (!(cls >= ST_MAP_MIN && map_mode_enabled() && map_pinned(cls, &p, &blk))) {
        cudaError_t e = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            // make room: drop the cache and retry once
            std::vector<std::pair<void *, PinnedBlk>> drop;
            {
                std::lock_guard<std::mutex> l(g_pool.mu);
                for (auto &kv : g_pool.free_blocks) drop.push_back(kv.second);
                g_pool.free_blocks.clear();
                g_pool.cached_bytes = 0;
            }
            for (auto &q : drop) release_pinned(q.first, q.second);
            cudaGetLastError();
            e = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            st_set_error("st_host_alloc: cudaHostAlloc(%zu) failed: %s", cls, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? ST_ERR_NOMEM : ST_ERR_CUDA;
        }
    }
    std::lock_guard<std::mutex> l(g_pool.mu);
    g_pool.live[p] = blk;
    *out = p;
    return ST_OK;
}

extern "C" int st_host_alloc(int64_t bytes, void **out) {
    if (!out || bytes < 0) return ST_ERR_INVALID_ARG;
    return st_pinned_alloc(size_t(bytes), out);
}

extern "C" int st_host_free(void *p) {
    if (!p) return ST_OK;
    PinnedBlk blk;
    bool keep = false;
    {
        std::lock_guard<std::mutex> l(g_pool.mu);
        auto it = g_pool.live.find(p);
        if (it == g_pool.live.end()) {
            st_set_error("st_host_free: pointer was not allocated by st_host_alloc");
            return ST_ERR_INVALID_ARG;
        }
        blk = it->second;
        g_pool.live.erase(it);
        if (!g_pool.cap_read) {
            g_pool.cap_read = true;
            if (const char *e = getenv("SUCHTREE_B200_PINNED_CACHE_MB")) g_pool.cap_bytes = size_t(atoll(e)) << 20;
        }
        if (g_pool.cached_bytes + blk.cls <= g_pool.cap_bytes) {
            g_pool.free_blocks.emplace(blk.cls, std::make_pair(p, blk));
            g_pool.cached_bytes += blk.cls;
            keep = true;
        }
    }
    if (!keep) release_pinned(p, blk);
    return ST_OK;
}

extern "C" int st_host_trim(int64_t keep_bytes) {
    std::vector<std::pair<void *, PinnedBlk>> drop;
    {
        std::lock_guard<std::mutex> l(g_pool.mu);
        while (!g_pool.free_blocks.empty() && int64_t(g_pool.cached_bytes) > std::max<int64_t>(keep_bytes, 0)) {
            auto it = std::prev(g_pool.free_blocks.end());  // largest first
            drop.push_back(it->second);
            g_pool.cached_bytes -= it->first;
            g_pool.free_blocks.erase(it);
        }
    }
    for (auto &q : drop) release_pinned(q.first, q.second);
    return ST_OK;
}

// ------------------------------------------------- caller-buffer registration -
// Page-locks a caller's (pageable) array in place so the host pipeline can DMA straight
// out of it -- no packing pass, no staging copy.  The Python shim does this for large
// inputs it sees repeatedly and unregisters when the array is garbage-collected.
extern "C" int st_host_register(const void *p, int64_t bytes) {
    if (!p || bytes <= 0) return ST_ERR_INVALID_ARG;
    cudaError_t e = cudaHostRegister(const_cast<void *>(p), size_t(bytes), cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return ST_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_host_register: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}

extern "C" int st_host_unregister(const void *p) {
    if (!p) return ST_OK;
    if (cudaHostUnregister(const_cast<void *>(p)) != cudaSuccess) cudaGetLastError();
    return ST_OK;
}

extern "C" int st_host_is_pinned(const void *p) { return st_is_pinned(p) ? 1 : 0; }

// ---------------------------------------------------------- copy roofline ----
// What the host interface can carry: `iters` rounds of an H2D copy of h2d_bytes and a
// D2H copy of d2h_bytes, pinned host memory, two streams (both directions at once),
// in chunks of chunk_bytes (the host pipeline's chunking).  *seconds = wall time per
// round, measured between two host-side synchronisations.
extern "C" int st_bench_copy(int device, int64_t h2d_bytes, int64_t d2h_bytes, int64_t chunk_bytes, int iters,
                             double *seconds) {
    if (!seconds || h2d_bytes < 0 || d2h_bytes < 0 || iters < 1 || (h2d_bytes == 0 && d2h_bytes == 0)) {
        st_set_error("st_bench_copy: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    if (chunk_bytes <= 0) chunk_bytes = int64_t(64) << 20;
    DeviceGuard g(device);
    if (!g.ok) {
        st_set_error("st_bench_copy: cudaSetDevice(%d) failed", device);
        return ST_ERR_CUDA;
    }
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    auto cleanup = [&]() {
        if (h_in) st_host_free(h_in);
        if (h_out) st_host_free(h_out);
        cudaFree(d_in);
        cudaFree(d_out);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    };
    int rc = ST_OK;
    if (h2d_bytes && (rc = st_host_alloc(h2d_bytes, &h_in)) != ST_OK) { cleanup(); return rc; }
    if (d2h_bytes && (rc = st_host_alloc(d2h_bytes, &h_out)) != ST_OK) { cleanup(); return rc; }
    if ((h2d_bytes && cudaMalloc(&d_in, size_t(h2d_bytes)) != cudaSuccess) ||
        (d2h_bytes && cudaMalloc(&d_out, size_t(d2h_bytes)) != cudaSuccess) ||
        cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        st_set_error("st_bench_copy: allocation failed");
        return ST_ERR_NOMEM;
    }
    if (h_in) memset(h_in, 1, size_t(h2d_bytes));  // touch: first-touch placement, no lazy zero pages
    if (h_out) memset(h_out, 1, size_t(d2h_bytes));
    auto round = [&]() {
        for (int64_t o = 0; o < std::max(h2d_bytes, d2h_bytes); o += chunk_bytes) {
            if (o < h2d_bytes)
                cudaMemcpyAsync(static_cast<char *>(d_in) + o, static_cast<char *>(h_in) + o,
                                size_t(std::min(chunk_bytes, h2d_bytes - o)), cudaMemcpyHostToDevice, s_in);
            if (o < d2h_bytes)
                cudaMemcpyAsync(static_cast<char *>(h_out) + o, static_cast<char *>(d_out) + o,
                                size_t(std::min(chunk_bytes, d2h_bytes - o)), cudaMemcpyDeviceToHost, s_out);
        }
    };
    round();  // warm
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_out);
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) round();
    cudaError_t e0 = cudaStreamSynchronize(s_in), e1 = cudaStreamSynchronize(s_out);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cleanup();
    if (e0 != cudaSuccess || e1 != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_bench_copy: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
        return ST_ERR_CUDA;
    }
    *seconds = dt / iters;
    return ST_OK;
}
