// Batched quartet topologies.
//
// Replaces SuchTree._quartet_topologies (MuchTree.pyx:1331-1376): six _mrca calls
// per quartet (each an O(depth^2) pointer chase in the reference) become THREE
// range-minimum depth lookups over the four (sorted) endpoint records -- 4 random
// sectors + 64 B streamed (int64x4 in, int64x4 out; 32 B with int32 ids on the device)
// per quartet.  The selection rule is the reference's: C[j] = how many of the six
// MRCAs equal M[j]; the first j with C[j] == 1 picks row j of the permutation table I
// (:1319-1320); no such j -> row 5 -- with MRCA equality decided from depths (below).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "st_device.cuh"
#include "st_hostctx.cuh"
#include "st_hostpool.cuh"

#ifndef ST_QPT_DEFAULT
#define ST_QPT_DEFAULT 1
#endif

// The reference's rule needs only the EQUALITY PATTERN of the six pair MRCAs, and that
// follows from their DEPTHS alone: two MRCAs that share an endpoint lie on that endpoint's
// root path, so they are the same node iff their depths are equal; for the three splits
// into disjoint pairs, M(p,q) == M(r,s) iff the depths are equal and M(p,r) is at least as
// deep (then both are the node at that depth on the common part of p's and r's root paths).
// With the four ids sorted (x0 <= x1 <= x2 <= x3, in-order ranks), every pair's MRCA depth is
// a minimum of the three ADJACENT range minima d1 = D(x0,x1), d2 = D(x1,x2), d3 = D(x2,x3):
// three range-minimum lookups instead of six, no MRCA ids, no root distances -- 8 of the
// record's 16 bytes (compact layout) per endpoint.
template <typename IdxT>
struct QuadT {
    IdxT v[4];
};

template <typename IdxT>
__device__ __forceinline__ QuadT<IdxT> quad_load(const IdxT *p, bool aligned) {
    QuadT<IdxT> q;
    if (aligned) {
        if (sizeof(IdxT) == 8) {
            uint64_t w[4];
            st_ld_stream_256(p, w);
#pragma unroll
            for (int k = 0; k < 4; ++k) q.v[k] = IdxT(w[k]);
        } else {
            const int4 w = st_ld_stream_int4(p);
            q.v[0] = IdxT(w.x); q.v[1] = IdxT(w.y); q.v[2] = IdxT(w.z); q.v[3] = IdxT(w.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) q.v[k] = __ldg(p + k);
    }
    return q;
}
template <typename IdxT>
__device__ __forceinline__ void quad_store(IdxT *p, bool aligned, IdxT a, IdxT b, IdxT c, IdxT d) {
    if (aligned) {
        if (sizeof(IdxT) == 8)
            asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                         "l"((long long)a), "l"((long long)b), "l"((long long)c), "l"((long long)d),
                         "l"(st_policy_evict_first())
                         : "memory");
        else
            st_st_stream_i32x4(reinterpret_cast<int32_t *>(p), int32_t(a), int32_t(b), int32_t(c), int32_t(d));
    } else {
        p[0] = a; p[1] = b; p[2] = c; p[3] = d;
    }
}

// (depth of the suffix argmin, depth of the prefix argmin) of one node: the second half
// of its record -- one 8-byte (compact) or 16-byte (wide) load, <= one sector
struct RecKeys {
    uint32_t suf_depth, pre_depth;
};
template <int M>
__device__ __forceinline__ RecKeys st_ld_keys(const TreeView &tv, int32_t id) {
    RecKeys r;
    if (st_compact<M>(tv)) {
        uint32_t sf, pr;
        asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];"
                     : "=r"(sf), "=r"(pr)
                     : "l"(reinterpret_cast<const char *>(tv.rec16 + id) + 8));
        r.suf_depth = sf >> tv.block_shift;
        r.pre_depth = pr >> tv.block_shift;
    } else {
        uint64_t sf, pr;
        asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];"
                     : "=l"(sf), "=l"(pr)
                     : "l"(reinterpret_cast<const char *>(tv.rec + id) + 16));
        r.suf_depth = uint32_t(sf >> 32);
        r.pre_depth = uint32_t(pr >> 32);
    }
    return r;
}

// depth of the minimum-depth node among ids [lo, hi] (lo <= hi): the depth of MRCA(lo, hi)
template <int M>
__device__ __forceinline__ uint32_t st_rmq_depth(const TreeView &tv, const SmemTables &sm, int32_t lo,
                                                 int32_t hi, uint32_t suf_lo, uint32_t pre_hi) {
    if (lo == hi) return uint32_t(__ldg(tv.depth + lo));  // repeated id: MRCA(a,a) = a
    const int32_t blo = lo >> tv.block_shift, bhi = hi >> tv.block_shift;
    if (blo == bhi)
        return uint32_t(st_rmq_inblock(tv.depth, tv.mst, tv.n_micro, tv.micro_shift, lo, hi) >> 32);
    uint32_t best = min(suf_lo, pre_hi);
    const int32_t span = bhi - blo - 1;
    if (span > 0) {
        const int k = 31 - __clz(span);
        if (st_ctab<M>(tv)) {
            const uint32_t *lvl = sm.stk32 + k * tv.n_blocks;
            best = min(best, min(lvl[blo + 1], lvl[bhi - (1 << k)]) >> tv.table_shift);
        } else {
            const uint64_t *lvl = sm.stk + k * tv.n_blocks;
            best = min(best, uint32_t(st_min64(lvl[blo + 1], lvl[bhi - (1 << k)]) >> 32));
        }
    }
    return best;
}

// index of the pair (u, v), u != v, in the reference's order ab ac ad bc bd cd
__device__ __forceinline__ int quartet_pair_index(int u, int v) {
    const int a = min(u, v), b = max(u, v);
    return 2 * a + b - 1 - (a == 2);
}

// P quartets per thread per iteration: 4P independent record gathers in flight
// SMALL (n_nodes <= 2^29): the sort runs on packed 32-bit keys (id << 2 | input position), two
// min/max instructions per comparator.
// PF: the ids of the thread's NEXT iteration are fetched before the current one is worked on.
// IdxT: type of the ids read, OutT: type of the ids written (the host path ships int32 ids
// over PCIe and gets the drop-in's int64 rows back).
template <int M, typename IdxT, int P, int MINB, bool SMALL, bool PF = false, typename OutT = IdxT, int QQT = 256>
__global__ void __launch_bounds__(QQT, MINB)
k_quartets(const TreeView tv, const IdxT *__restrict__ quartets, int64_t n, OutT *__restrict__ out,
           int aligned) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar;
    const SmemTables sm = st_load_tables<M>(tv, smem_raw, &tables_bar);
    const long long nn = tv.n_nodes;
    const int64_t groups = (n + P - 1) / P;
    const int64_t stride = int64_t(gridDim.x) * QQT;
    auto fetch = [&](int64_t g, QuadT<IdxT>(&dst)[P]) {
#pragma unroll
        for (int t = 0; t < P; ++t) {
            const int64_t i = g * P + t;
            if (g < groups && i < n) dst[t] = quad_load<IdxT>(quartets + 4 * i, aligned != 0);
            else dst[t].v[0] = dst[t].v[1] = dst[t].v[2] = dst[t].v[3] = 0;
        }
    };
    QuadT<IdxT> nxt[P];
    if (PF) fetch(int64_t(blockIdx.x) * QQT + threadIdx.x, nxt);
    for (int64_t g = int64_t(blockIdx.x) * QQT + threadIdx.x; g < groups; g += stride) {
        QuadT<IdxT> q[P];
        int32_t x[P][4];  // ids sorted ascending
        int pos[P][4];    // pos[r] = position in the input quartet of the id of rank r
        bool ok[P];
        RecKeys rk[P][4];
        if (PF) {
#pragma unroll
            for (int t = 0; t < P; ++t) q[t] = nxt[t];
            fetch(g + stride, nxt);
        } else {
            fetch(g, q);
        }
#pragma unroll
        for (int t = 0; t < P; ++t) {
            const int64_t i = g * P + t;
            ok[t] = i < n;
            // range check of quartet_topologies_bulk (MuchTree.pyx:1303-1310), on the device
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) bad |= (unsigned long long)(long long)q[t].v[k] >= (unsigned long long)nn;
            if (bad) {
                long long mx = q[t].v[0], mn = q[t].v[0];
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    mx = (long long)q[t].v[k] > mx ? (long long)q[t].v[k] : mx;
                    mn = (long long)q[t].v[k] < mn ? (long long)q[t].v[k] : mn;
                }
                if (mx >= nn) atomicMax(&tv.status->max_bad, (unsigned long long)mx);
                if (mn < 0) atomicMin(&tv.status->min_bad, mn);
                if (ok[t]) quad_store<OutT>(out + 4 * i, aligned != 0, OutT(-1), OutT(-1), OutT(-1), OutT(-1));
                ok[t] = false;
            }
            if (SMALL) {
                // 5-comparator sorting network on packed keys
                int32_t key[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) key[k] = ok[t] ? (int32_t(q[t].v[k]) << 2) | k : k;
#define ST_CE(a, b)                                          \
    {                                                        \
        const int32_t lo_ = min(key[a], key[b]);             \
        key[b] = max(key[a], key[b]);                        \
        key[a] = lo_;                                        \
    }
                ST_CE(0, 1) ST_CE(2, 3) ST_CE(0, 2) ST_CE(1, 3) ST_CE(1, 2)
#undef ST_CE
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    x[t][k] = key[k] >> 2;
                    pos[t][k] = key[k] & 3;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    x[t][k] = ok[t] ? int32_t(q[t].v[k]) : 0;
                    pos[t][k] = k;
                }
                // 5-comparator sorting network on (id, position)
#define ST_CE(a, b)                                          \
    {                                                        \
        const bool sw = x[t][a] > x[t][b];                   \
        const int32_t xa = x[t][a], xb = x[t][b];            \
        const int pa = pos[t][a], pb = pos[t][b];            \
        x[t][a] = sw ? xb : xa; x[t][b] = sw ? xa : xb;      \
        pos[t][a] = sw ? pb : pa; pos[t][b] = sw ? pa : pb;  \
    }
                ST_CE(0, 1) ST_CE(2, 3) ST_CE(0, 2) ST_CE(1, 3) ST_CE(1, 2)
#undef ST_CE
            }
        }
#pragma unroll
        for (int t = 0; t < P; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) rk[t][k] = st_ld_keys<M>(tv, x[t][k]);
#pragma unroll
        for (int t = 0; t < P; ++t) {
            if (!ok[t]) continue;
            const uint32_t d1 = st_rmq_depth<M>(tv, sm, x[t][0], x[t][1], rk[t][0].suf_depth, rk[t][1].pre_depth);
            const uint32_t d2 = st_rmq_depth<M>(tv, sm, x[t][1], x[t][2], rk[t][1].suf_depth, rk[t][2].pre_depth);
            const uint32_t d3 = st_rmq_depth<M>(tv, sm, x[t][2], x[t][3], rk[t][2].suf_depth, rk[t][3].pre_depth);
            // MRCA depths of the six pairs of RANKS: 01 02 03 12 13 23
            const uint32_t D01 = d1, D12 = d2, D23 = d3, D02 = min(d1, d2), D13 = min(d2, d3), D03 = min(D02, d3);
            // C[k] = how many of the six MRCAs equal pair k's (itself included).  Pairs sharing a
            // rank: equal depths.  Disjoint splits: 01|23 needs d2 >= d1 besides (M(x0,x2) at least as
            // deep as M(x0,x1)); 02|13 and 03|12: the linking MRCA M(x0,x1) is never shallower.
            const int e01_02 = D01 == D02, e01_03 = D01 == D03, e01_12 = D01 == D12, e01_13 = D01 == D13;
            const int e01_23 = (D01 == D23) & (d2 >= d1);
            const int e02_03 = D02 == D03, e02_12 = D02 == D12, e02_13 = D02 == D13, e02_23 = D02 == D23;
            const int e03_12 = D03 == D12, e03_13 = D03 == D13, e03_23 = D03 == D23;
            const int e12_13 = D12 == D13, e12_23 = D12 == D23, e13_23 = D13 == D23;
            const int c01 = 1 + e01_02 + e01_03 + e01_12 + e01_13 + e01_23;
            const int c02 = 1 + e01_02 + e02_03 + e02_12 + e02_13 + e02_23;
            const int c03 = 1 + e01_03 + e02_03 + e03_12 + e03_13 + e03_23;
            const int c12 = 1 + e01_12 + e02_12 + e03_12 + e12_13 + e12_23;
            const int c13 = 1 + e01_13 + e02_13 + e03_13 + e12_13 + e13_23;
            const int c23 = 1 + e01_23 + e02_23 + e03_23 + e12_23 + e13_23;
            // the reference scans the pairs of input POSITIONS in the order ab ac ad bc bd cd and
            // takes the first whose MRCA is unique; none -> the fall-through row 5
            int j = 5;
            j = min(j, c01 == 1 ? quartet_pair_index(pos[t][0], pos[t][1]) : 5);
            j = min(j, c02 == 1 ? quartet_pair_index(pos[t][0], pos[t][2]) : 5);
            j = min(j, c03 == 1 ? quartet_pair_index(pos[t][0], pos[t][3]) : 5);
            j = min(j, c12 == 1 ? quartet_pair_index(pos[t][1], pos[t][2]) : 5);
            j = min(j, c13 == 1 ? quartet_pair_index(pos[t][1], pos[t][3]) : 5);
            j = min(j, c23 == 1 ? quartet_pair_index(pos[t][2], pos[t][3]) : 5);
            // rows of I: {0,1,2,3} {0,2,1,3} {0,3,1,2} {1,2,0,3} {1,3,0,2} {2,3,0,1}
            const OutT a = OutT(q[t].v[0]), b = OutT(q[t].v[1]), c = OutT(q[t].v[2]), d = OutT(q[t].v[3]);
            OutT o0, o1, o2, o3;
            switch (j) {
                case 0: o0 = a; o1 = b; o2 = c; o3 = d; break;
                case 1: o0 = a; o1 = c; o2 = b; o3 = d; break;
                case 2: o0 = a; o1 = d; o2 = b; o3 = c; break;
                case 3: o0 = b; o1 = c; o2 = a; o3 = d; break;
                case 4: o0 = b; o1 = d; o2 = a; o3 = c; break;
                default: o0 = c; o1 = d; o2 = a; o3 = b; break;
            }
            quad_store<OutT>(out + 4 * (g * P + t), aligned != 0, o0, o1, o2, o3);
        }
    }
}

static int st_quartets_per_thread() {  // SUCHTREE_B200_QPT = 1 | 2 (read per launch: tests flip it)
    const char *e = getenv("SUCHTREE_B200_QPT");
    const int x = e ? atoi(e) : ST_QPT_DEFAULT;
    return (x == 1 || x == 2) ? x : ST_QPT_DEFAULT;
}

template <int M, typename IdxT, int P, int MINB, bool SMALL, bool PF = false, typename OutT = IdxT, int QQT = 256>
static int launch_quartets_p(const st_tree *t, const IdxT *d_q, int64_t n, OutT *d_out,
                             cudaStream_t stream, RangeStatus *status) {
    auto kern = k_quartets<M, IdxT, P, MINB, SMALL, PF, OutT, QQT>;
    const int smem = t->query_smem_bytes;
    int rc = st_raise_smem(kern, t->device, smem);
    if (rc != ST_OK) return rc;
    int per_sm = 0;
    ST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QQT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t groups = (n + P - 1) / P;
    const int grid = int(std::min<int64_t>((groups + QQT - 1) / QQT, int64_t(t->sm_count) * per_sm));
    const int aligned = (reinterpret_cast<uintptr_t>(d_q) % (4 * sizeof(IdxT)) == 0) &&
                        (reinterpret_cast<uintptr_t>(d_out) % (4 * sizeof(OutT)) == 0);
    TreeView view = t->view;
    if (status) view.status = status;
    kern<<<grid, QQT, smem, stream>>>(view, d_q, n, d_out, aligned);
    ST_CUDA(cudaGetLastError());
    return ST_OK;
}

template <int M, typename IdxT>
static int launch_quartets_m(const st_tree *t, const IdxT *d_q, int64_t n, IdxT *d_out,
                             cudaStream_t stream, RangeStatus *status) {
    if (t->n_nodes > (int64_t(1) << 29)) return launch_quartets_p<M, IdxT, 1, 4, false>(t, d_q, n, d_out, stream, status);
    if (st_quartets_per_thread() == 2) return launch_quartets_p<M, IdxT, 2, 3, true>(t, d_q, n, d_out, stream, status);
    // int64 ids (the drop-in layout): 3 CTAs of 384 threads at 56 registers, 1152 threads per SM,
    // measured 4.26e10 vs 4.15e10 quartets/s for 4 x 256 at 64; int32 ids: 4 x 256 is the faster (4.61e10 vs 4.52e10)
    if (sizeof(IdxT) == 8) return launch_quartets_p<M, IdxT, 1, 3, true, false, IdxT, 384>(t, d_q, n, d_out, stream, status);
    // (measured and not kept, profiles/r02_summary.md: 5 resident CTAs per SM -- 48 registers, spills,
    //  0.59x; prefetching the next iteration's ids, PF = true, 0.94x; two quartets per thread 0.91x)
    return launch_quartets_p<M, IdxT, 1, 4, true>(t, d_q, n, d_out, stream, status);
}

template <typename IdxT>
static int launch_quartets(const st_tree *t, const IdxT *d_q, int64_t n, IdxT *d_out,
                           cudaStream_t stream, RangeStatus *status = nullptr) {
    if (n == 0) return ST_OK;
    if (t->compact) return launch_quartets_m<1, IdxT>(t, d_q, n, d_out, stream, status);
    if (t->compact_tables) return launch_quartets_m<3, IdxT>(t, d_q, n, d_out, stream, status);
    return launch_quartets_m<0, IdxT>(t, d_q, n, d_out, stream, status);
}

// int32 ids in, int64 rows out (host path)
static int launch_quartets_mixed(const st_tree *t, const int32_t *d_q, int64_t n, int64_t *d_out,
                                 cudaStream_t stream, RangeStatus *status) {
    if (n == 0) return ST_OK;
    if (t->n_nodes > (int64_t(1) << 29)) {
        if (t->compact) return launch_quartets_p<1, int32_t, 1, 4, false, false, int64_t>(t, d_q, n, d_out, stream, status);
        if (t->compact_tables) return launch_quartets_p<3, int32_t, 1, 4, false, false, int64_t>(t, d_q, n, d_out, stream, status);
        return launch_quartets_p<0, int32_t, 1, 4, false, false, int64_t>(t, d_q, n, d_out, stream, status);
    }
    if (t->compact) return launch_quartets_p<1, int32_t, 1, 4, true, false, int64_t>(t, d_q, n, d_out, stream, status);
    if (t->compact_tables) return launch_quartets_p<3, int32_t, 1, 4, true, false, int64_t>(t, d_q, n, d_out, stream, status);
    return launch_quartets_p<0, int32_t, 1, 4, true, false, int64_t>(t, d_q, n, d_out, stream, status);
}

extern "C" int st_quartet_topologies_device(const st_tree *t, const int64_t *d_quartets, int64_t n,
                                            int64_t *d_out, void *stream) {
    if (!t || n < 0 || (n > 0 && (!d_quartets || !d_out))) {
        st_set_error("st_quartet_topologies_device: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(t->device);
    return launch_quartets<int64_t>(t, d_quartets, n, d_out, static_cast<cudaStream_t>(stream));
}

extern "C" int st_quartet_topologies_device32(const st_tree *t, const int32_t *d_quartets, int64_t n,
                                              int32_t *d_out, void *stream) {
    if (!t || n < 0 || (n > 0 && (!d_quartets || !d_out))) {
        st_set_error("st_quartet_topologies_device32: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(t->device);
    return launch_quartets<int32_t>(t, d_quartets, n, d_out, static_cast<cudaStream_t>(stream));
}

// host buffers: the same 3-slot pipeline as st_distances on a lane of the device's host
// context -- pack(chunk c+1: int64 ids -> int32 in pinned staging, host pool) | H2D (16 B per
// quartet) + kernel + D2H (int64 rows, 32 B per quartet) of chunk c | copy-out(chunk c-2).
// Rows land straight in a page-locked `out` (the shim's result arrays), else in pinned staging
// first.  Small calls: one kernel over the pinned mappings (zero-copy), one synchronisation.
extern "C" int st_quartet_topologies(const st_tree *t, const int64_t *quartets, int64_t s0, int64_t s1,
                                     int64_t n, int64_t *out) {
    if (!t || n < 0 || (n > 0 && (!quartets || !out))) {
        st_set_error("st_quartet_topologies: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    if (n == 0) return ST_OK;
    DeviceGuard g(t->device);
    LaneGuard lg(t->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    auto range_error = [&](unsigned long long mxb, long long mnb) {
        // the reference reports max_id when it is >= size, else min_id (MuchTree.pyx:1303-1310)
        st_set_bad_node(mxb != 0 ? (int64_t)mxb : (int64_t)mnb);
        st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(), (long long)t->n_nodes);
        return ST_ERR_NODE_RANGE;
    };
    unsigned long long mxb = 0;
    long long mnb = 0;
    int rc = ST_OK;
    if (n <= ST_MEDIUM_CALL / 4) {  // h_small_in / h_small_out hold ST_MEDIUM_CALL pairs = 1/4 as many int64 rows
        int32_t *hp = static_cast<int32_t *>(lane->h_small_in);
        if (st_pack_ids(quartets, s0, s1, n, hp, 4) >> 31) {
            st_report_range(quartets, s0, s1, n, 4, t->n_nodes);
            return ST_ERR_NODE_RANGE;
        }
        int64_t *ho = static_cast<int64_t *>(lane->h_small_out);
        cudaStream_t st = lane->streams[0];
        rc = launch_quartets_mixed(t, hp, n, ho, st, lane->d_status);
        if (rc != ST_OK) return rc;
        rc = st_lane_read_status(lane, st, &mxb, &mnb);  // synchronises
        if (rc != ST_OK) return rc;
        if (mxb != 0 || mnb != 0) return range_error(mxb, mnb);
        memcpy(out, ho, size_t(n) * 32);
        return ST_OK;
    }
    const bool out_pinned = st_is_pinned(out) && st_is_pinned(out + 4 * n - 1);
    rc = st_lane_ensure_stage(lane, 4 * n, true, !out_pinned);
    if (rc != ST_OK) return rc;
    const int64_t C = lane->stage_pairs / 4;  // rows per chunk: d_out / h_out hold 8 B per staged pair
    int64_t chunk_begin[ST_LANE_SLOTS] = {0, 0, 0}, chunk_len[ST_LANE_SLOTS] = {0, 0, 0};
    auto copy_out = [&](int s) {
        if (!out_pinned && chunk_len[s] > 0)
            st_parallel_copy(out + 4 * chunk_begin[s], lane->h_out[s], size_t(chunk_len[s]) * 32);
        chunk_len[s] = 0;
    };
    auto quiesce = [&]() {
        for (int k = 0; k < ST_LANE_SLOTS; ++k) cudaStreamSynchronize(lane->streams[k]);
        st_lane_read_status(lane, lane->streams[0], &mxb, &mnb);
    };
    int64_t done = 0;
    int c = 0;
    for (; done < n; ++c) {
        const int s = c % ST_LANE_SLOTS;
        const int64_t m = std::min(C, n - done);
        cudaStream_t st = lane->streams[s];
        if (c >= ST_LANE_SLOTS) {
            ST_CUDA(cudaEventSynchronize(lane->ev[s]));
            copy_out(s);
        }
        int32_t *hp = static_cast<int32_t *>(lane->h_in[s]);
        if (st_pack_ids(quartets + done * s0, s0, s1, m, hp, 4) >> 31) {
            quiesce();
            st_report_range(quartets, s0, s1, n, 4, t->n_nodes);
            return ST_ERR_NODE_RANGE;
        }
        int32_t *d_in = static_cast<int32_t *>(lane->d_in[s]);
        int64_t *d_o = static_cast<int64_t *>(lane->d_out[s]);
        ST_CUDA(cudaMemcpyAsync(d_in, hp, size_t(m) * 16, cudaMemcpyHostToDevice, st));
        rc = launch_quartets_mixed(t, d_in, m, d_o, st, lane->d_status);
        if (rc != ST_OK) {
            quiesce();
            return rc;
        }
        void *dst = out_pinned ? static_cast<void *>(out + 4 * done) : lane->h_out[s];
        ST_CUDA(cudaMemcpyAsync(dst, d_o, size_t(m) * 32, cudaMemcpyDeviceToHost, st));
        ST_CUDA(cudaEventRecord(lane->ev[s], st));
        chunk_begin[s] = done;
        chunk_len[s] = m;
        done += m;
    }
    for (int k = 0; k < ST_LANE_SLOTS && k < c; ++k) {
        const int s = (c - std::min(c, ST_LANE_SLOTS) + k) % ST_LANE_SLOTS;
        ST_CUDA(cudaStreamSynchronize(lane->streams[s]));
        copy_out(s);
    }
    rc = st_lane_read_status(lane, lane->streams[0], &mxb, &mnb);
    if (rc != ST_OK) return rc;
    if (mxb != 0 || mnb != 0) return range_error(mxb, mnb);
    return ST_OK;
}
