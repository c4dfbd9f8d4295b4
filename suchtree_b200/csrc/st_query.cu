// Batched (a,b) -> (MRCA, patristic distance) queries.
//
// Replaces SuchTree._mrca (MuchTree.pyx:999-1030) and SuchTree._distances
// (MuchTree.pyx:911-943): instead of recording a's ancestor chain and scanning it
// for every ancestor of b (O(depth^2) pointer chasing per pair), the MRCA is the
// argmin of depth over the id interval [min(a,b), max(a,b)] (ids are in-order
// ranks), answered in O(1) from
//     rec[lo].suf, rec[hi].pre            one record each: 16 or 32 bytes, <= one sector,
//                                         and they also carry rd[lo], rd[hi]
//     two block-table entries             shared memory (staged by one TMA bulk copy)
// and the distance is rd[a] + rd[b] - 2 rd[mrca] in double-double, with rd[mrca] from
// the shared-memory table of block minima (a third gather only when the MRCA is not a
// block minimum: ~3 % of far-apart pairs).
// => 2 random L2 sectors + 16 B (int32x2 in, fp64 out) of streamed HBM per pair.
// The same file holds the host-buffer pipeline of the drop-in entry points.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>
#include "st_device.cuh"
#ifndef ST_LEAN_DEFAULT
#define ST_LEAN_DEFAULT 1
#endif
#include "st_hostctx.cuh"
#include "st_hostpool.cuh"

// P pairs per thread per iteration, fetched as raw 64-bit words by 16-byte (P = 2,
// int32) or 32-byte streaming loads and decoded when used
template <typename IdxT, int P>
struct RawPairs {
    static constexpr int W = P * int(sizeof(IdxT)) / 4;  // 64-bit words
    uint64_t w[W];
};
template <typename IdxT, int P>
__device__ __forceinline__ RawPairs<IdxT, P> st_load_pairs(const IdxT *pairs, int64_t first_pair) {
    RawPairs<IdxT, P> r;
    const IdxT *p = pairs + 2 * first_pair;
    if (P == 1) {
        if (sizeof(IdxT) == 4) {
            r.w[0] = uint64_t(uint32_t(__ldg(p))) | (uint64_t(uint32_t(__ldg(p + 1))) << 32);
        } else {
            r.w[0] = uint64_t(__ldg(reinterpret_cast<const long long *>(p)));
            r.w[RawPairs<IdxT, P>::W - 1] = uint64_t(__ldg(reinterpret_cast<const long long *>(p) + 1));
        }
    } else if (RawPairs<IdxT, P>::W == 2) {
        const int4 v = st_ld_stream_int4(p);
        r.w[0] = uint64_t(uint32_t(v.x)) | (uint64_t(uint32_t(v.y)) << 32);
        r.w[1] = uint64_t(uint32_t(v.z)) | (uint64_t(uint32_t(v.w)) << 32);
    } else {
#pragma unroll
        for (int h = 0; h < RawPairs<IdxT, P>::W / 4; ++h) {
            uint64_t q[4];
            st_ld_stream_256(reinterpret_cast<const char *>(p) + 32 * h, q);
#pragma unroll
            for (int k = 0; k < 4; ++k) r.w[4 * h + k] = q[k];
        }
    }
    return r;
}
template <typename IdxT, int P>
__device__ __forceinline__ void st_decode_pair(const RawPairs<IdxT, P> &r, int k, long long &a, long long &b) {
    if (sizeof(IdxT) == 4) {
        a = int32_t(uint32_t(r.w[k]));
        b = int32_t(uint32_t(r.w[k] >> 32));
    } else {
        a = (long long)r.w[2 * k];
        b = (long long)r.w[2 * k + 1];
    }
}

// Bit-packed pair stream of the host pipeline (st_pack_pairs_bits): pair i = bits [i * 2w, (i+1) * 2w)
// of a little-endian bit stream, w = tv.id_bits; a in the low w bits, b above it.  Three aligned
// 32-bit words cover any pair (2w <= 62, bit offset in the first word <= 31); neighbouring lanes
// read overlapping words, which L1 serves.  The buffer has one spare 64-bit word at its end.
struct StPackedPairs {};
__device__ __forceinline__ void st_load_pair_packed(const uint32_t *__restrict__ base, int64_t i, int w,
                                                    long long &a, long long &b) {
    const uint64_t off = uint64_t(i) * uint64_t(2 * w);
    const uint32_t *p = base + (off >> 5);
    const int sh = int(off & 31);
    const uint32_t x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2);
    const uint64_t lo = uint64_t(x0) | (uint64_t(x1) << 32);
    const uint64_t v = (lo >> sh) | (sh ? (uint64_t(x2) << (64 - sh)) : 0ull);
    const uint64_t m = (uint64_t(1) << w) - 1;
    a = (long long)(v & m);
    b = (long long)((v >> w) & m);
}

// range check as SuchTree.distances_bulk does before the kernel (MuchTree.pyx:897-903),
// moved onto the device so the host never makes a pass over the pair array.
__device__ __forceinline__ PairQ st_make_query(const TreeView &tv, long long a, long long b) {
    PairQ q;
    const long long n = tv.n_nodes;
    q.bad = (unsigned long long)a >= (unsigned long long)n || (unsigned long long)b >= (unsigned long long)n;
    if (q.bad) {
        long long mx = a > b ? a : b, mn = a < b ? a : b;
        if (mx >= n) atomicMax(&tv.status->max_bad, (unsigned long long)mx);
        if (mn < 0) atomicMin(&tv.status->min_bad, mn);
        a = b = 0;
    }
    int32_t x = int32_t(a), y = int32_t(b);
    q.lo = min(x, y);
    q.hi = max(x, y);
    return q;
}

// One pair: both endpoint records, the block table, the distance and/or the MRCA id.
template <int M>
__device__ __forceinline__ void st_pair(const TreeView &tv, const SmemTables &sm, const PairQ &q,
                                        const RecRaw &l, const RecRaw &h, bool want_d, bool want_m,
                                        double &d, int32_t &m) {
    bool ft = false;
    // MRCA(a,a) = a
    const uint64_t k = q.lo == q.hi ? uint64_t(uint32_t(q.lo))
                                    : st_rmq<M>(tv, sm, q.lo, q.hi, l.suf, h.pre, &ft);
    if (want_d) d = st_patristic(dd{l.rd_hi, l.rd_lo}, dd{h.rd_hi, h.rd_lo}, st_mrca_rd<M>(tv, sm, k, ft));
    if (want_m) m = st_mrca_id<M>(tv, sm, k, ft);
}

// P pairs per thread per iteration: 2P independent record gathers are in flight
// before anything depends on them (the kernel is latency-bound on those gathers).
// PR selects the form of the per-pair work:
//   0  generic (any layout: wide records, double-double root distances, 64-bit keys)
//   2  lean compact (st_pair_c: 3 registers per record, 32-bit keys) -- the default for compact trees
//   3  lean compact with PAIRED records: each endpoint's whole 32-byte sector is fetched with one
//      256-bit load, and rd[mrca] is taken from an endpoint or its sector neighbour when the MRCA
//      is one (ladder-like trees; chosen per tree by the build-time probe, st_tree_create)
template <typename IdxT, int P, int M, int QT, int MINB, int PR>
__global__ void __launch_bounds__(QT, MINB)
k_pairs(const TreeView tv, const IdxT *__restrict__ pairs, int64_t n, double *__restrict__ out,
        int32_t *__restrict__ mrca_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar;
    const SmemTables sm = st_load_tables<M>(tv, smem_raw, &tables_bar);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const bool want_d = out != nullptr, want_m = mrca_out != nullptr;

    const int64_t groups = n / P;
    const int64_t stride = int64_t(gridDim.x) * QT;
    int64_t i = int64_t(blockIdx.x) * QT + threadIdx.x;
    // (prefetching the next iteration's ids, and 4 pairs per thread, were both measured
    //  and are no faster: profiles/r01_summary.md)
    for (; i < groups; i += stride) {
        PairQ q[P];
        if constexpr (std::is_same<IdxT, StPackedPairs>::value) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
                long long a, b;
                st_load_pair_packed(reinterpret_cast<const uint32_t *>(pairs), P * i + k, tv.id_bits, a, b);
                q[k] = st_make_query(tv, a, b);
            }
        } else {
            const RawPairs<IdxT, P> cur = st_load_pairs<IdxT, P>(pairs, P * i);
#pragma unroll
            for (int k = 0; k < P; ++k) {
                long long a, b;
                st_decode_pair<IdxT, P>(cur, k, a, b);
                q[k] = st_make_query(tv, a, b);
            }
        }
        if constexpr (PR >= 2) {  // lean compact path (PR = 3: with the sector neighbours)
            constexpr bool NB = PR == 3;
            RecC lc[P], hc[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                lc[k] = st_ld_rec_c<NB>(tv, q[k].lo, false);
                hc[k] = st_ld_rec_c<NB>(tv, q[k].hi, true);
            }
            double dc[P];
            int32_t mc[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                dc[k] = 0.0;
                mc[k] = 0;
                st_pair_c<NB>(tv, sm, q[k], lc[k], hc[k], want_d, want_m, dc[k], mc[k]);
                if (q[k].bad) {
                    dc[k] = nan;
                    mc[k] = -1;
                }
            }
            if (want_d) {
                if (P == 4) st_st_stream_f64x4(out + P * i, dc[0], dc[1], dc[2], dc[3]);
                else if (P == 2) st_st_stream_f64x2(out + P * i, dc[0], dc[1]);
                else st_st_stream_f64(out + i, dc[0]);
            }
            if (want_m) {
                if (P == 4) st_st_stream_i32x4(mrca_out + P * i, mc[0], mc[1], mc[2], mc[3]);
                else if (P == 2) st_st_stream_i32x2(mrca_out + P * i, mc[0], mc[1]);
                else st_st_stream_i32(mrca_out + i, mc[0]);
            }
        } else {
            RecRaw l[P], h[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                l[k] = st_ld_rec<M>(tv, q[k].lo);
                h[k] = st_ld_rec<M>(tv, q[k].hi);
            }
            double d[P];
            int32_t m[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                d[k] = 0.0;
                m[k] = 0;
                st_pair<M>(tv, sm, q[k], l[k], h[k], want_d, want_m, d[k], m[k]);
                if (q[k].bad) {
                    d[k] = nan;
                    m[k] = -1;
                }
            }
            if (want_d) {
                if (P == 4) st_st_stream_f64x4(out + P * i, d[0], d[1], d[2], d[3]);
                else if (P == 2) st_st_stream_f64x2(out + P * i, d[0], d[1]);
                else st_st_stream_f64(out + i, d[0]);
            }
            if (want_m) {
                if (P == 4) st_st_stream_i32x4(mrca_out + P * i, m[0], m[1], m[2], m[3]);
                else if (P == 2) st_st_stream_i32x2(mrca_out + P * i, m[0], m[1]);
                else st_st_stream_i32(mrca_out + i, m[0]);
            }
        }
    }
    // the n % P pairs left over: one thread each
    const int64_t t = int64_t(blockIdx.x) * QT + threadIdx.x;
    if (P > 1 && t < n - groups * P) {
        const int64_t i = groups * P + t;
        long long a, b;
        if constexpr (std::is_same<IdxT, StPackedPairs>::value)
            st_load_pair_packed(reinterpret_cast<const uint32_t *>(pairs), i, tv.id_bits, a, b);
        else
            st_decode_pair<IdxT, 1>(st_load_pairs<IdxT, 1>(pairs, i), 0, a, b);
        const PairQ q = st_make_query(tv, a, b);
        const RecRaw l = st_ld_rec<M>(tv, q.lo), h = st_ld_rec<M>(tv, q.hi);
        double d = 0.0;
        int32_t m = 0;
        st_pair<M>(tv, sm, q, l, h, want_d, want_m, d, m);
        if (want_d) out[i] = q.bad ? nan : d;
        if (want_m) mrca_out[i] = q.bad ? -1 : m;
    }
}

// ------------------------------------------------------------------ launch --
template <typename IdxT, int P, int M, int PR = 0, int QT_ = 0, int MINB_ = 0>
static int launch_variant_m(const st_tree *t, const void *d_pairs, int64_t n, double *d_out,
                            int32_t *d_mrca, cudaStream_t stream, RangeStatus *status) {
    // 2 pairs/thread: 2 x 512 threads x 64 registers; 4 pairs/thread needs ~80 registers:
    // 3 x 256 threads (8 gathers in flight per thread)
#ifndef ST_QT_P2
#define ST_QT_P2 512
#define ST_MINB_P2 2
#endif
    constexpr int QT = QT_ ? QT_ : (P == 4 ? 256 : ST_QT_P2), MINB = MINB_ ? MINB_ : (P == 4 ? 3 : ST_MINB_P2);
    auto kern = k_pairs<IdxT, P, M, QT, MINB, PR>;
    // per device, per thread: the attribute and the occupancy query are made once per
    // shared-memory size (they cost microseconds that single-pair calls would notice)
    static thread_local int configured_smem[64] = {0};
    static thread_local int cached_smem[64] = {0}, cached_per_sm[64] = {0};
    const int smem = t->query_smem_bytes, dv = t->device & 63;
    if (smem > 48 * 1024 && configured_smem[dv] < smem) {
        // raise-only and process-wide (st_hostctx.cuh): a thread with smaller tables can never
        // lower the limit under another thread's launch; the thread-local mark only skips the lock
        const int rc = st_raise_smem(kern, t->device, smem);
        if (rc != ST_OK) return rc;
        configured_smem[dv] = smem;
    }
    if (cached_per_sm[dv] == 0 || cached_smem[dv] != smem) {
        int q = 0;
        ST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, QT, smem));
        cached_per_sm[dv] = q < 1 ? 1 : q;
        cached_smem[dv] = smem;
    }
    const int per_sm = cached_per_sm[dv];
    const int64_t items = std::max<int64_t>(n / P, 1);
    int64_t want = (items + QT - 1) / QT;
    int grid = int(std::min<int64_t>(want, int64_t(t->sm_count) * per_sm));
    if (grid < 1) grid = 1;
    TreeView view = t->view;
    if (status) view.status = status;  // host calls bring their own status word (st_hostctx.cuh)
    kern<<<grid, QT, smem, stream>>>(view, static_cast<const IdxT *>(d_pairs), n, d_out, d_mrca);
    ST_CUDA(cudaGetLastError());
    return ST_OK;
}

// the tree's own choice (build-time probe, st_tree_create) unless SUCHTREE_B200_PAIRED = 0 | 1
// forces it (read per launch: tests and experiments flip it)
static int st_paired_records(const st_tree *t) {
    const char *e = getenv("SUCHTREE_B200_PAIRED");
    return e && e[0] ? (atoi(e) != 0) : t->paired;
}

template <typename IdxT, int P>
static int launch_variant(const st_tree *t, const void *d_pairs, int64_t n, double *d_out,
                          int32_t *d_mrca, cudaStream_t stream, RangeStatus *status) {
    if (t->compact && P == 2 && st_paired_records(t))
        return launch_variant_m<IdxT, 2, 1, 3>(t, d_pairs, n, d_out, d_mrca, stream, status);
    if (t->compact) {
        // lean compact path (3 registers per record, 32-bit keys): +1 % over the generic one on
        // every tree shape, bit-identical; SUCHTREE_B200_LEAN = 0 keeps the generic path (tests).
        // Measured and not kept (profiles/r02_lean_variants.json): 384 x 3 (equal), 256 x 5 and
        // 512 x 3 (spills: 0.55x / 0.77x), four pairs per thread at 256 x 3 (0.94x) or 512 x 2 (0.79x).
        const char *e = getenv("SUCHTREE_B200_LEAN");
        const int lean = e && e[0] ? atoi(e) : ST_LEAN_DEFAULT;
        if (lean == 0) return launch_variant_m<IdxT, P, 1>(t, d_pairs, n, d_out, d_mrca, stream, status);
        if constexpr (P == 4) return launch_variant_m<IdxT, 4, 1, 2, 256, 3>(t, d_pairs, n, d_out, d_mrca, stream, status);
        return launch_variant_m<IdxT, P, 1, 2>(t, d_pairs, n, d_out, d_mrca, stream, status);
    }
    if (t->compact_tables) return launch_variant_m<IdxT, P, 3>(t, d_pairs, n, d_out, d_mrca, stream, status);
    return launch_variant_m<IdxT, P, 0>(t, d_pairs, n, d_out, d_mrca, stream, status);
}

static int st_pairs_per_thread() {  // SUCHTREE_B200_PPT = 1 | 2 | 4 (read per launch: experiments)
    const char *e = getenv("SUCHTREE_B200_PPT");
    const int x = e && e[0] ? atoi(e) : 2;
    return (x == 1 || x == 2 || x == 4) ? x : 2;
}

int st_launch_pairs(const st_tree *t, const void *d_pairs, int idx_bits, int64_t n, double *d_out,
                    int32_t *d_mrca, cudaStream_t stream, RangeStatus *status) {
    if (n == 0) return ST_OK;
    if (idx_bits == ST_IDX_PACKED) {  // the host pipeline's bit-packed stream (2 x tv.id_bits per pair, 4-byte aligned)
        const bool ok16 = (!d_out || reinterpret_cast<uintptr_t>(d_out) % 16 == 0) &&
                          (!d_mrca || reinterpret_cast<uintptr_t>(d_mrca) % 8 == 0);
        if (ok16) return launch_variant<StPackedPairs, 2>(t, d_pairs, n, d_out, d_mrca, stream, status);
        return launch_variant<StPackedPairs, 1>(t, d_pairs, n, d_out, d_mrca, stream, status);
    }
    if (idx_bits != 32 && idx_bits != 64) {
        st_set_error("idx_bits must be 32 or 64 (got %d)", idx_bits);
        return ST_ERR_INVALID_ARG;
    }
    auto aligned = [&](uintptr_t in_al, uintptr_t d_al, uintptr_t m_al) {
        return (reinterpret_cast<uintptr_t>(d_pairs) % in_al == 0) &&
               (!d_out || reinterpret_cast<uintptr_t>(d_out) % d_al == 0) &&
               (!d_mrca || reinterpret_cast<uintptr_t>(d_mrca) % m_al == 0);
    };
    const int ppt = st_pairs_per_thread();
    if (idx_bits == 32) {
        if (ppt >= 4 && aligned(32, 32, 16)) return launch_variant<int32_t, 4>(t, d_pairs, n, d_out, d_mrca, stream, status);
        if (ppt >= 2 && aligned(16, 16, 8)) return launch_variant<int32_t, 2>(t, d_pairs, n, d_out, d_mrca, stream, status);
        return launch_variant<int32_t, 1>(t, d_pairs, n, d_out, d_mrca, stream, status);
    }
    if (ppt >= 4 && aligned(32, 32, 16)) return launch_variant<int64_t, 4>(t, d_pairs, n, d_out, d_mrca, stream, status);
    if (ppt >= 2 && aligned(32, 16, 8)) return launch_variant<int64_t, 2>(t, d_pairs, n, d_out, d_mrca, stream, status);
    return launch_variant<int64_t, 1>(t, d_pairs, n, d_out, d_mrca, stream, status);
}

int st_read_range_status(const st_tree *t, cudaStream_t stream, bool *bad) {
    RangeStatus h{};
    ST_CUDA(cudaMemcpyAsync(&h, t->d_status, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ST_CUDA(cudaStreamSynchronize(stream));
    *bad = false;
    if (h.max_bad != 0 || h.min_bad != 0) {
        *bad = true;
        // the reference reports max_id when it is >= size, else min_id (MuchTree.pyx:899-903)
        st_set_bad_node(h.max_bad != 0 ? (int64_t)h.max_bad : (int64_t)h.min_bad);
        ST_CUDA(cudaMemsetAsync(t->d_status, 0, sizeof(RangeStatus), stream));
        ST_CUDA(cudaStreamSynchronize(stream));
        st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(),
                     (long long)t->n_nodes);
    }
    return ST_OK;
}

#ifndef ST_PAIRS_KERNEL_ID
#define ST_PAIRS_KERNEL_ID "unknown"
#endif
extern "C" const char *st_pairs_kernel_id(void) { return ST_PAIRS_KERNEL_ID; }

// ------------------------------------------------------------ device API ----
extern "C" int st_distances_device(const st_tree *t, const void *d_pairs, int idx_bits, int64_t n,
                                   double *d_out, int32_t *d_mrca, void *stream) {
    if (!t || n < 0 || (n > 0 && !d_pairs) || (!d_out && !d_mrca && n > 0)) {
        st_set_error("st_distances_device: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    if (idx_bits != 32 && idx_bits != 64) {
        st_set_error("idx_bits must be 32 or 64 (got %d)", idx_bits);
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(t->device);
    return st_launch_pairs(t, d_pairs, idx_bits, n, d_out, d_mrca, static_cast<cudaStream_t>(stream));
}

extern "C" int st_check_range(const st_tree *t, void *stream) {
    if (!t) return ST_ERR_INVALID_ARG;
    DeviceGuard g(t->device);
    bool bad = false;
    int rc = st_read_range_status(t, static_cast<cudaStream_t>(stream), &bad);
    if (rc != ST_OK) return rc;
    return bad ? ST_ERR_NODE_RANGE : ST_OK;
}

// ---------------------------------------------------------- synthetic input -
template <typename IdxT>
__global__ void k_random_leaf_pairs(uint32_t n_leaves, uint64_t seed, int64_t first, int64_t n,
                                    IdxT *__restrict__ pairs) {
    // one Philox call yields two pairs
    const int64_t n2 = (n + 1) >> 1;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n2;
         i += int64_t(gridDim.x) * blockDim.x) {
        // counter is the index of the pair-of-pairs in the GLOBAL stream, so a
        // shard starting at an even `first` reproduces the same numbers
        Philox4 r = st_philox4x32_10(uint64_t((first >> 1) + i), seed);
        int64_t p = 2 * i;
        pairs[2 * p] = IdxT(2 * st_bounded(r.x, n_leaves));
        pairs[2 * p + 1] = IdxT(2 * st_bounded(r.y, n_leaves));
        if (p + 1 < n) {
            pairs[2 * p + 2] = IdxT(2 * st_bounded(r.z, n_leaves));
            pairs[2 * p + 3] = IdxT(2 * st_bounded(r.w, n_leaves));
        }
    }
}

extern "C" int st_random_leaf_pairs_device(const st_tree *t, uint64_t seed, int64_t first_pair,
                                           int64_t n, void *d_pairs, int idx_bits, void *stream) {
    if (!t || n < 0 || (n > 0 && !d_pairs) || (first_pair & 1)) {
        st_set_error("st_random_leaf_pairs_device: bad arguments (first_pair must be even)");
        return ST_ERR_INVALID_ARG;
    }
    if (n == 0) return ST_OK;
    DeviceGuard g(t->device);
    const int TPB = 256;
    int64_t want = ((n + 1) / 2 + TPB - 1) / TPB;
    int grid = int(std::min<int64_t>(want, int64_t(t->sm_count) * 16));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (idx_bits == 32)
        k_random_leaf_pairs<int32_t><<<grid, TPB, 0, s>>>(uint32_t(t->n_leaves), seed, first_pair, n,
                                                          static_cast<int32_t *>(d_pairs));
    else if (idx_bits == 64)
        k_random_leaf_pairs<int64_t><<<grid, TPB, 0, s>>>(uint32_t(t->n_leaves), seed, first_pair, n,
                                                          static_cast<int64_t *>(d_pairs));
    else {
        st_set_error("idx_bits must be 32 or 64");
        return ST_ERR_INVALID_ARG;
    }
    ST_CUDA(cudaGetLastError());
    return ST_OK;
}

// -------------------------------------------------------------- host API ----
// Chunked 3-slot pipeline on a LANE of the device's host context (st_hostctx.cuh):
// pack(chunk c+1) on the host pool | H2D + kernel + D2H of chunk c on one of the lane's
// three streams | copy-out(chunk c-2).  The caller's int64 ids (the drop-in dtype, any
// strides, pageable or pinned) are BIT-PACKED into pinned staging by the host thread pool:
// 2 x ceil(log2(n_nodes)) bits per pair (4.5 bytes for a 100k-leaf tree instead of 16), which
// the pair kernel decodes itself (StPackedPairs).  What bounds the call on one GPU is the host's
// memory system, and the staging write + DMA read is the part of that traffic the width decides.  Results go straight to
// the caller's buffer when it is page-locked (the Python shim's result arrays come from
// the pinned pool, st_host_alloc), else through pinned staging + a parallel copy.
// A pinned (or registered) contiguous input is partly DMA'd as it is and read by the
// int64 kernel variant (no host pass over that part).
// Every call owns its lane's range-status word: concurrent callers on one tree neither
// queue on a mutex nor see each other's out-of-range flags.

// Latency path for small and medium calls (distance(a,b), common_ancestor(a,b), lists up to
// 2^18 pairs): the ids are packed on the host into pinned staging, ONE kernel reads them and
// writes the results through the pinned mappings (zero-copy over PCIe), one synchronisation.
static int host_pairs_small(const st_tree *t, HostLane *lane, const int64_t *pairs, int64_t s0, int64_t s1,
                            int64_t n, double *out_d, int32_t *out_m) {
    int rc = ST_OK;
    int32_t *hp = static_cast<int32_t *>(lane->h_small_in);
    const bool tiny = n <= ST_SMALL_CALL;
    if (tiny) {
        int64_t mx = INT64_MIN, mn = INT64_MAX;
        for (int64_t i = 0; i < n; ++i) {
            const int64_t a = pairs[i * s0], b = pairs[i * s0 + s1];
            mx = std::max(mx, std::max(a, b));
            mn = std::min(mn, std::min(a, b));
            hp[2 * i] = int32_t(a);
            hp[2 * i + 1] = int32_t(b);
        }
        if (mn < 0 || mx >= t->n_nodes) {  // the reference's report (MuchTree.pyx:897-903)
            st_set_bad_node(mx >= t->n_nodes ? mx : mn);
            st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(),
                         (long long)t->n_nodes);
            return ST_ERR_NODE_RANGE;
        }
    } else if (st_pack_ids(pairs, s0, s1, n, hp, 2) >> 31) {  // negative or beyond int32
        st_report_range(pairs, s0, s1, n, 2, t->n_nodes);
        return ST_ERR_NODE_RANGE;
    }
    void *ho = lane->h_small_out;
    cudaStream_t st = lane->streams[0];
    rc = st_launch_pairs(t, hp, 32, n, out_d ? static_cast<double *>(ho) : nullptr,
                         out_m ? static_cast<int32_t *>(ho) : nullptr, st, lane->d_status);
    if (rc != ST_OK) return rc;
    if (tiny) {
        ST_CUDA(cudaStreamSynchronize(st));  // ids were checked on the host: the status word stays clear
    } else {
        // ids in [n_nodes, 2^31) are caught by the kernel: the lane's status word rides the same stream
        unsigned long long mxb = 0;
        long long mnb = 0;
        rc = st_lane_read_status(lane, st, &mxb, &mnb);
        if (rc != ST_OK) return rc;
        if (mxb != 0 || mnb != 0) {
            st_report_range(pairs, s0, s1, n, 2, t->n_nodes);
            return ST_ERR_NODE_RANGE;
        }
    }
    const size_t bytes = size_t(n) * (out_d ? 8 : 4);
    void *dst = out_d ? static_cast<void *>(out_d) : static_cast<void *>(out_m);
    if (tiny) memcpy(dst, ho, bytes);
    else st_parallel_copy(dst, ho, bytes);
    return ST_OK;
}

// fraction of every chunk of a page-locked input whose ids are bit-packed by the host (the rest is
// DMA'd as int64); see host_pairs_run
static double st_default_pack_fraction() {
    double pack_fraction = 0.7;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
        // two ranks, fraction 0 / 0.2 / 0.4 / 0.6: 6.10 / 6.35 / 6.12 / 5.75e9 pairs/s in total
        const int lw = atoi(e);
        if (lw == 2) pack_fraction = 0.2;
        else if (lw > 2) pack_fraction = 0.0;
    }
    if (const char *e = getenv("SUCHTREE_B200_PACK_FRACTION")) pack_fraction = std::min(1.0, std::max(0.0, atof(e)));
    return pack_fraction;
}

// Pairs per chunk of the pipeline: the staging slot for long calls; a mid-size call is cut
// into several chunks so that packing, the two copy directions and the kernel overlap within it
// (SUCHTREE_B200_CHUNK_PAIRS overrides: experiments).
static int64_t st_chunk_pairs(int64_t n, int64_t stage_pairs) {
    // about four chunks per call, 2^18 .. stage_pairs pairs each (scripts/midsize_exp.py: 4e6 pairs
    // 1.72 -> 1.14 ms, 2e6 0.82 -> 0.66 ms, 1e6 0.45 -> 0.39 ms; 3e7 and beyond: whole slots)
    int64_t c = int64_t(1) << 18;
    while (c * 8 <= n) c <<= 1;
    c = std::min(c, stage_pairs);
    if (const char *e = getenv("SUCHTREE_B200_CHUNK_PAIRS")) {
        const int64_t v = atoll(e);
        if (v >= 4096) c = std::min(stage_pairs, v & ~int64_t(3));
    }
    return c;
}

static int host_pairs_run(const st_tree *t, const int64_t *pairs, int64_t s0, int64_t s1, int64_t n,
                          double *out_d, int32_t *out_m) {
    if (!t || n < 0 || (n > 0 && (!pairs || (!out_d == !out_m)))) {  // exactly one output
        st_set_error("bad arguments (NULL pointer or negative n)");
        return ST_ERR_INVALID_ARG;
    }
    if (n == 0) return ST_OK;
    DeviceGuard g(t->device);
    LaneGuard lg(t->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    if (n <= ST_MEDIUM_CALL) return host_pairs_small(t, lane, pairs, s0, s1, n, out_d, out_m);
    const bool contiguous = (s1 == 1 && s0 == 2);
    const bool out_pinned = st_is_pinned(out_d ? static_cast<void *>(out_d) : static_cast<void *>(out_m));
    const bool in_pinned = contiguous && st_is_pinned(pairs) && st_is_pinned(pairs + 2 * n - 1);
    bool pack = true;
    if (in_pinned && st_host_threads() < 4) pack = false;
    if (const char *e = getenv("SUCHTREE_B200_HOST_PATH")) {
        if (e[0] == 'd' && in_pinned) pack = false;  // "direct"
        if (e[0] == 'p') pack = true;                // "pack"
    }
    // hybrid: PCIe wants the ids packed (4.5 instead of 16 B/pair for a 100k-leaf tree), host memory
    // bandwidth wants them left alone (packing costs a CPU read of the ids plus the staging write and
    // its DMA read on top of what the DMA of the results moves): pack a fraction of every chunk and
    // ship the rest as int64.  One GPU, 1e8 pairs, fraction 0 / 0.3 / 0.45 / 0.6 / 0.8 / 1:
    // 3.14 / 3.91 / 4.30 / 4.64 / 4.67 / 4.18e9 pairs/s, 0.7: 4.90e9 (profiles/r02_packfrac_bits*.json;
    // with the int32 packing of earlier rounds the curve was flat at 3.7-3.8e9).  With several ranks on one
    // host (one process per GPU) the host memory system is the shared bottleneck and plain DMA wins
    // (8-GPU box, int32 packing: 6.3e9 at 0 % vs 5.4e9 at 45 %).
    const double pack_fraction = st_default_pack_fraction();
    const bool hybrid = pack && pack_fraction < 1.0 && in_pinned;
    int rc = st_lane_ensure_stage(lane, n, pack, !out_pinned);
    if (rc != ST_OK) return rc;
    const int64_t C = st_chunk_pairs(n, lane->stage_pairs);
    const size_t out_elem = out_d ? 8 : 4;
    char *user_out = out_d ? reinterpret_cast<char *>(out_d) : reinterpret_cast<char *>(out_m);
    int64_t chunk_begin[ST_LANE_SLOTS] = {0, 0, 0}, chunk_len[ST_LANE_SLOTS] = {0, 0, 0};
    auto copy_out = [&](int s) {  // results of the chunk last issued on slot s -> caller's buffer
        if (!out_pinned && chunk_len[s] > 0)
            st_parallel_copy(user_out + size_t(chunk_begin[s]) * out_elem, lane->h_out[s],
                          size_t(chunk_len[s]) * out_elem);
        chunk_len[s] = 0;
    };
    auto quiesce = [&]() {  // error exits: nothing of this call may still be in flight on the lane
        for (int k = 0; k < ST_LANE_SLOTS; ++k) cudaStreamSynchronize(lane->streams[k]);
        unsigned long long mxb = 0;
        long long mnb = 0;
        st_lane_read_status(lane, lane->streams[0], &mxb, &mnb);  // clears anything the kernels flagged
    };
    int64_t done = 0;
    int c = 0;
    for (; done < n; ++c) {
        const int s = c % ST_LANE_SLOTS;
        const int64_t m = std::min(C, n - done);
        cudaStream_t st = lane->streams[s];
        if (c >= ST_LANE_SLOTS) {
            ST_CUDA(cudaEventSynchronize(lane->ev[s]));  // slot s: device buffers and staging are free again
            copy_out(s);
        }
        const int64_t *src = pairs + done * s0;
        double *dd_out = out_d ? static_cast<double *>(lane->d_out[s]) : nullptr;
        int32_t *dm_out = out_m ? static_cast<int32_t *>(lane->d_out2[s]) : nullptr;
        // first mp pairs: bit-packed by the host pool; the rest (hybrid mode, pinned
        // contiguous input): DMA'd as int64 while the pool packs
        const int64_t mp = pack ? (hybrid ? (int64_t(double(m) * pack_fraction) & ~int64_t(3)) : m) : 0;
        char *d_in = static_cast<char *>(lane->d_in[s]);
        // packed part: the bit-packed pair stream (2 x id_bits per pair + one spare word); the
        // unpacked part follows it, 32-byte aligned
        const int w = t->view.id_bits;
        const size_t packed_bytes = mp > 0 ? ((size_t(mp) * size_t(2 * w) + 63) / 64 + 1) * 8 : 0;
        char *d_direct = d_in + ((packed_bytes + 31) & ~size_t(31));
        if (mp < m)
            ST_CUDA(cudaMemcpyAsync(d_direct, src + 2 * mp, size_t(m - mp) * 16, cudaMemcpyHostToDevice, st));
        if (mp > 0) {
            uint64_t *hp = static_cast<uint64_t *>(lane->h_in[s]);
            const uint64_t acc = st_pack_pairs_bits(src, s0, s1, mp, hp, w);
            if (acc >> w) {  // a negative id, or one of more than id_bits bits: not a node of this tree
                quiesce();
                st_report_range(pairs, s0, s1, n, 2, t->n_nodes);
                return ST_ERR_NODE_RANGE;
            }
            ST_CUDA(cudaMemcpyAsync(d_in, hp, packed_bytes, cudaMemcpyHostToDevice, st));
            rc = st_launch_pairs(t, d_in, ST_IDX_PACKED, mp, dd_out, dm_out, st, lane->d_status);
            if (rc != ST_OK) {
                quiesce();
                return rc;
            }
        }
        if (mp < m) {
            rc = st_launch_pairs(t, d_direct, 64, m - mp, dd_out ? dd_out + mp : nullptr,
                                 dm_out ? dm_out + mp : nullptr, st, lane->d_status);
            if (rc != ST_OK) {
                quiesce();
                return rc;
            }
        }
        void *dst = out_pinned ? static_cast<void *>(user_out + size_t(done) * out_elem) : lane->h_out[s];
        const void *dsrc = out_d ? static_cast<const void *>(dd_out) : static_cast<const void *>(dm_out);
        ST_CUDA(cudaMemcpyAsync(dst, dsrc, size_t(m) * out_elem, cudaMemcpyDeviceToHost, st));
        ST_CUDA(cudaEventRecord(lane->ev[s], st));
        chunk_begin[s] = done;
        chunk_len[s] = m;
        done += m;
    }
    // drain in issue order
    for (int k = 0; k < ST_LANE_SLOTS && k < c; ++k) {
        const int s = (c - std::min(c, ST_LANE_SLOTS) + k) % ST_LANE_SLOTS;
        ST_CUDA(cudaStreamSynchronize(lane->streams[s]));
        copy_out(s);
    }
    unsigned long long mxb = 0;
    long long mnb = 0;
    rc = st_lane_read_status(lane, lane->streams[0], &mxb, &mnb);
    if (rc != ST_OK) return rc;
    if (mxb != 0 || mnb != 0) {
        // the reference reports max_id when it is >= size, else min_id (MuchTree.pyx:899-903)
        st_set_bad_node(mxb != 0 ? (int64_t)mxb : (int64_t)mnb);
        st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(), (long long)t->n_nodes);
        return ST_ERR_NODE_RANGE;
    }
    return ST_OK;
}

// measurement helper: rate of the host-side id packing alone (pinned src and dst)
extern "C" int st_bench_pack(int64_t n_pairs, int iters, double *pairs_per_s) {
    if (!pairs_per_s || n_pairs < 1 || iters < 1) return ST_ERR_INVALID_ARG;
    int64_t *src = nullptr;
    int32_t *dst = nullptr;
    ST_CUDA(cudaMallocHost(&src, size_t(n_pairs) * 16));
    if (cudaMallocHost(&dst, size_t(n_pairs) * 8) != cudaSuccess) {
        cudaFreeHost(src);
        return ST_ERR_NOMEM;
    }
    for (int64_t i = 0; i < 2 * n_pairs; ++i) src[i] = i & 0xffff;
    uint64_t acc = st_pack_ids(src, 2, 1, n_pairs, dst, 2);  // warm (page touch)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaEventSynchronize(e0);
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) acc |= st_pack_ids(src, 2, 1, n_pairs, dst, 2);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeHost(src);
    cudaFreeHost(dst);
    *pairs_per_s = (acc >> 31) ? 0.0 : double(n_pairs) * iters / dt;
    return ST_OK;
}

extern "C" int st_host_route_info(const st_tree *t, double *pack_fraction, int *id_bits) {
    if (!t) return ST_ERR_INVALID_ARG;
    if (pack_fraction) *pack_fraction = st_default_pack_fraction();
    if (id_bits) *id_bits = t->view.id_bits;
    return ST_OK;
}

extern "C" int st_host_pack_pairs(const int64_t *pairs, int64_t stride0, int64_t stride1, int64_t n, int id_bits,
                                  uint64_t *out, uint64_t *or_of_ids) {
    if (!pairs || !out || n < 0 || id_bits < 1 || id_bits > 31) {
        st_set_error("st_host_pack_pairs: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    const uint64_t acc = st_pack_pairs_bits(pairs, stride0, stride1, n, out, id_bits);
    if (or_of_ids) *or_of_ids = acc;
    return ST_OK;
}

extern "C" int st_distances(const st_tree *t, const int64_t *pairs, int64_t stride0, int64_t stride1,
                            int64_t n, double *out) {
    return host_pairs_run(t, pairs, stride0, stride1, n, out, nullptr);
}

extern "C" int st_mrca(const st_tree *t, const int64_t *pairs, int64_t stride0, int64_t stride1,
                       int64_t n, int32_t *out) {
    return host_pairs_run(t, pairs, stride0, stride1, n, nullptr, out);
}
