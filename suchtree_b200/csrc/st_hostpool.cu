// Host thread pool (see st_hostpool.cuh).  Plain C++: no device code here.
#include "st_hostpool.cuh"

#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "st_internal.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace {

struct Pool {
    std::mutex mu;                 // protects the fields below
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int, int)> *fn = nullptr;
    int n_parts = 0, next_part = 0, pending = 0;
    uint64_t generation = 0;
    bool stop = false;
    std::mutex call_mu;            // one parallel_for at a time
    int n_threads = 1;

    Pool() {
        int hw = int(std::thread::hardware_concurrency());
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = std::min(hw > 0 ? hw : 1 << 20, CPU_COUNT(&set));
#endif
        if (hw < 1) hw = 1;
        // one process per GPU (torchrun): share the cores between the local ranks
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
            const int lw = atoi(e);
            if (lw > 1) hw = std::max(1, hw / lw);
        }
        n_threads = std::min(hw, 16);
        if (const char *e = getenv("SUCHTREE_B200_HOST_THREADS")) {
            int v = atoi(e);
            if (v >= 1 && v <= 256) n_threads = v;
        }
        for (int i = 1; i < n_threads; ++i) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
    }
    void run_parts(std::unique_lock<std::mutex> &l) {
        while (next_part < n_parts) {
            const int p = next_part++;
            const auto *f = fn;
            const int np = n_parts;
            l.unlock();
            (*f)(p, np);
            l.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    void loop() {
        std::unique_lock<std::mutex> l(mu);
        uint64_t seen = 0;
        for (;;) {
            cv_work.wait(l, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            run_parts(l);
        }
    }
    void parallel_for(int parts, const std::function<void(int, int)> &f) {
        if (parts <= 1 || n_threads <= 1) {
            for (int p = 0; p < parts; ++p) f(p, parts);
            return;
        }
        std::lock_guard<std::mutex> call(call_mu);
        std::unique_lock<std::mutex> l(mu);
        fn = &f;
        n_parts = parts;
        next_part = 0;
        pending = parts;
        ++generation;
        cv_work.notify_all();
        run_parts(l);  // the caller works too
        cv_done.wait(l, [&] { return pending == 0; });
        fn = nullptr;
    }
};

Pool &pool() {
    static Pool *p = new Pool();  // leaked on purpose: no static-destruction order issues at exit
    return *p;
}

}  // namespace

void st_parallel_for(int n_parts, const std::function<void(int, int)> &fn) { pool().parallel_for(n_parts, fn); }
int st_host_threads() { return pool().n_threads; }

// ------------------------------------------------- host passes over id arrays --
// int64 ids -> int32, OR of everything seen (bit 31 and above set <=> some id is
// negative or >= 2^31: the rare error path then finds the exact id)
static uint64_t pack_ids_contig(const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t count) {
    uint64_t acc = 0;
    int64_t i = 0;
#if defined(__SSE2__)
    // 4 ids per step; streaming stores: the staging buffer is read next by the DMA
    // engine, not by this core, so skip the read-for-ownership of its lines
    while (i < count && (reinterpret_cast<uintptr_t>(dst + i) & 15)) {
        const int64_t v = src[i];
        acc |= uint64_t(v);
        dst[i++] = int32_t(v);
    }
    __m128i vacc = _mm_setzero_si128();
    for (; i + 8 <= count; i += 8) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 2));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 4));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 6));
        vacc = _mm_or_si128(vacc, _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d)));
        const __m128 lo = _mm_shuffle_ps(_mm_castsi128_ps(a), _mm_castsi128_ps(b), _MM_SHUFFLE(2, 0, 2, 0));
        const __m128 hi = _mm_shuffle_ps(_mm_castsi128_ps(c), _mm_castsi128_ps(d), _MM_SHUFFLE(2, 0, 2, 0));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_castps_si128(lo));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_castps_si128(hi));
    }
    alignas(16) uint64_t lanes[2];
    _mm_store_si128(reinterpret_cast<__m128i *>(lanes), vacc);
    acc |= lanes[0] | lanes[1];
    _mm_sfence();
#endif
    for (; i < count; ++i) {
        const int64_t v = src[i];
        acc |= uint64_t(v);
        dst[i] = int32_t(v);
    }
    return acc;
}
uint64_t st_pack_ids(const int64_t *src, int64_t s0, int64_t s1, int64_t m, int32_t *dst, int width) {
    const bool contiguous = (s1 == 1 && s0 == width);
    const int parts = int(std::min<int64_t>(st_host_threads(), m / 32768 + 1));
    std::vector<uint64_t> accs(size_t(parts), 0);
    st_parallel_for(parts, [&](int p, int np) {
        const int64_t b = m * p / np, e = m * (p + 1) / np;
        uint64_t acc = 0;
        if (contiguous) {
            acc = pack_ids_contig(src + b * width, dst + b * width, (e - b) * width);
        } else {
            for (int64_t i = b; i < e; ++i)
                for (int k = 0; k < width; ++k) {
                    const int64_t v = src[i * s0 + k * s1];
                    acc |= uint64_t(v);
                    dst[i * width + k] = int32_t(v);
                }
        }
        accs[size_t(p)] = acc;
    });
    uint64_t acc = 0;
    for (uint64_t a : accs) acc |= a;
    return acc;
}
// int64 id pairs -> bit-packed pairs: pair i occupies bits [i * 2w, (i + 1) * 2w) of a little-endian
// bit stream, a in its low w bits and b in the high w bits (w = id bits of the tree, 2w <= 62).
// 4.5 instead of 8 bytes per pair for a 100k-leaf tree: what the staging write and the DMA read
// cost the host's memory system.  Each part starts on a 64-pair boundary (= a 64-bit word
// boundary for every w) and streams whole 64-bit words (movnti).  Returns the OR of all ids: a
// bit at or above w set <=> some id is negative or >= 2^w (ids in [n_nodes, 2^w) are caught by
// the kernel).  dst needs room for ceil(m * 2w / 64) + 1 words.
static inline void st_stream_u64(uint64_t *p, uint64_t v) {
#if defined(__SSE2__) && defined(__x86_64__)
    _mm_stream_si64(reinterpret_cast<long long *>(p), (long long)v);
#else
    *p = v;
#endif
}
// One group of 64 contiguous pairs = exactly 2W words, every shift a compile-time constant
// (template recursion instead of a carried fill counter: straight-line code, no branches).
template <int W, int I>
struct PackStep {
    static __attribute__((always_inline)) inline void run(const int64_t *__restrict__ q, uint64_t &acc, uint64_t &seen,
                                                          uint64_t *__restrict__ &out) {
        constexpr int B = 2 * W, fill = (I * B) & 63;
        const uint64_t x = uint64_t(q[2 * I]), y = uint64_t(q[2 * I + 1]);
        seen |= x | y;
        const uint64_t v = (x | (y << W)) & ((uint64_t(1) << B) - 1);
        if constexpr (fill == 0) acc = v;
        else acc |= v << fill;
        if constexpr (fill + B >= 64) {
            st_stream_u64(out++, acc);
            constexpr int rest = fill + B - 64;  // bits of v that belong to the next word
            if constexpr (rest > 0) acc = v >> (B - rest);
            else acc = 0;
        }
        if constexpr (I + 1 < 64) PackStep<W, I + 1>::run(q, acc, seen, out);
    }
};
template <int W>
static uint64_t pack_groups(const int64_t *__restrict__ src, int64_t n_groups, uint64_t *__restrict__ out) {
    uint64_t seen = 0, acc = 0;
    for (int64_t g = 0; g < n_groups; ++g, src += 128) PackStep<W, 0>::run(src, acc, seen, out);
    return seen;
}
typedef uint64_t (*PackGroupsFn)(const int64_t *, int64_t, uint64_t *);
template <int... Ws>
static const PackGroupsFn *pack_group_table(std::integer_sequence<int, Ws...>) {
    static const PackGroupsFn table[] = {pack_groups<Ws + 1>...};  // W = 1 .. 31
    return table;
}

uint64_t st_pack_pairs_bits(const int64_t *src, int64_t s0, int64_t s1, int64_t m, uint64_t *dst, int w) {
    const int bits = 2 * w;
    const int64_t groups = (m + 63) / 64;
    const int parts = int(std::min<int64_t>(st_host_threads(), m / 32768 + 1));
    const bool contiguous = (s0 == 2 && s1 == 1);
    const PackGroupsFn fast = pack_group_table(std::make_integer_sequence<int, 31>())[w - 1];
    std::vector<uint64_t> accs(size_t(parts), 0);
    st_parallel_for(parts, [&](int p, int np) {
        int64_t b = (groups * p / np) * 64;
        const int64_t e = std::min<int64_t>(m, (groups * (p + 1) / np) * 64);
        uint64_t *out = dst + (b / 64) * bits;  // 64 pairs = `bits` words
        uint64_t seen = 0, acc = 0;
        if (contiguous && e - b >= 64) {  // whole groups: the unrolled form
            const int64_t ng = (e - b) / 64;
            seen = fast(src + 2 * b, ng, out);
            out += ng * bits;
            b += ng * 64;
        }
        int fill = 0;
        const int64_t *q = src + b * s0;
        for (int64_t i = b; i < e; ++i, q += s0) {  // strided input, and the last partial group
            const uint64_t x = uint64_t(q[0]), y = uint64_t(q[s1]);
            seen |= x | y;
            const uint64_t v = (x | (y << w)) & ((uint64_t(1) << bits) - 1);
            acc |= v << fill;
            fill += bits;
            if (fill >= 64) {
                fill -= 64;
                st_stream_u64(out++, acc);
                acc = fill ? v >> (bits - fill) : 0;
            }
        }
        if (fill) *out = acc;  // the last, partial word of the stream (only the last part can have one)
#if defined(__SSE2__)
        _mm_sfence();
#endif
        accs[size_t(p)] = seen;
    });
    uint64_t acc = 0;
    for (uint64_t a : accs) acc |= a;
    return acc;
}
void st_parallel_copy(void *dst, const void *src, size_t bytes) {
    const int parts = int(std::min<size_t>(size_t(st_host_threads()), bytes / (size_t(1) << 20) + 1));
    st_parallel_for(parts, [&](int p, int np) {
        const size_t b = bytes * size_t(p) / size_t(np), e = bytes * size_t(p + 1) / size_t(np);
        memcpy(static_cast<char *>(dst) + b, static_cast<const char *>(src) + b, e - b);
    });
}
// the reference's report for an out-of-range array: max id if it is >= size, else min id
void st_report_range(const int64_t *src, int64_t s0, int64_t s1, int64_t n, int width, int64_t n_nodes) {
    int64_t mx = INT64_MIN, mn = INT64_MAX;
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < width; ++k) {
            const int64_t v = src[i * s0 + k * s1];
            mx = v > mx ? v : mx;
            mn = v < mn ? v : mn;
        }
    st_set_bad_node(mx >= n_nodes ? mx : mn);
    st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(), (long long)n_nodes);
}

