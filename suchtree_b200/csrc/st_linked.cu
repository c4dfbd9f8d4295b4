// Linked-tree paths: exhaustive link-pair distances, the bucketed sampler with the
// reference's exact xorshift64* stream, and the Philox sampler with fused moments.
//
// Replaces SuchLinkedTrees.linked_distances (MuchTree.pyx:2900-2934),
// _random_int (:2936-2949) and the sampling loop of sample_linked_distances
// (:3023-3052).  Link pairs are generated ON the device from the link list: the
// (size,2) int64 id arrays the reference materialises are only written when the
// caller asks for them.
#include <algorithm>
#include <climits>
#include <cmath>
#include <mutex>
#include <vector>

#include <new>

#include "st_device.cuh"
#include "st_hostctx.cuh"
#include "st_hostpool.cuh"

static const int LT = 256;

// one full pair query (fast path tables in shared memory)
template <int M = 2>
__device__ __forceinline__ double linked_query(const TreeView &tv, const SmemTables &sm, int32_t a,
                                               int32_t b) {
    int32_t lo = min(a, b), hi = max(a, b);
    if (lo == hi) return 0.0;
    if (M == 1) {  // compact layout: the lean form (3 registers per record, 32-bit keys), same bits
        const RecC lc = st_ld_rec_c<false>(tv, lo, false), hc = st_ld_rec_c<false>(tv, hi, true);
        double d = 0.0;
        int32_t m = 0;
        st_pair_c<false>(tv, sm, PairQ{lo, hi, false}, lc, hc, true, false, d, m);
        return d;
    }
    RecRaw l = st_ld_rec<M>(tv, lo), h = st_ld_rec<M>(tv, hi);
    bool ft;
    uint64_t key = st_rmq<M>(tv, sm, lo, hi, l.suf, h.pre, &ft);
    return st_patristic(dd{l.rd_hi, l.rd_lo}, dd{h.rd_hi, h.rd_lo}, st_mrca_rd<M>(tv, sm, key, ft));
}

// Row-cached queries.  In the (i, j < i) enumeration consecutive pairs share their `i` link: its
// two records (one per tree) are loaded once per row and kept in registers, so a pair costs ONE
// new record gather per tree instead of two.  Compact layout only (M = 1); the other layouts
// take the plain query.
struct RecFull {
    double rd;
    uint32_t suf, pre;
};
__device__ __forceinline__ RecFull st_ld_rec_full(const TreeView &tv, int32_t id) {
    uint64_t a, b;
    asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(tv.rec16 + id));
    return RecFull{__longlong_as_double((long long)a), uint32_t(b), uint32_t(b >> 32)};
}
// distance(a, b) with a's record at hand
template <int M>
__device__ __forceinline__ double linked_query_row(const TreeView &tv, const SmemTables &sm, int32_t a,
                                                   const RecFull &ra, int32_t b) {
    if (M != 1) return linked_query<M>(tv, sm, a, b);
    if (a == b) return 0.0;
    const RecFull rb = st_ld_rec_full(tv, b);
    const bool a_lo = a < b;
    const RecC l{a_lo ? ra.rd : rb.rd, a_lo ? ra.suf : rb.suf, 0.0};
    const RecC h{a_lo ? rb.rd : ra.rd, a_lo ? rb.pre : ra.pre, 0.0};
    double d = 0.0;
    int32_t m = 0;
    st_pair_c<false>(tv, sm, PairQ{a_lo ? a : b, a_lo ? b : a, false}, l, h, true, false, d, m);
    return d;
}

// Joined link records.  A sample or a link pair needs, per link and per tree, the endpoint's root
// distance and its two block keys -- found so far by following the link row to one record per
// tree (three dependent random sectors per link).  When both trees have the compact layout
// and the fields fit, the handle keeps them JOINED per link: one 32-byte record = one sector
//     rd_a, rd_b                    root distances of the link's TreeA / TreeB leaf
//     pk_a, pk_b                    id | suffix key << ib | prefix key << (ib + kb), per tree
// so a sampled pair costs 2 random sectors instead of 6, and the exhaustive (i, j < i)
// enumeration -- lanes adjacent in j -- reads them CONTIGUOUSLY: no random gathers at all beyond
// the ~3 % of pairs whose MRCA is not a block minimum.  Same arithmetic (st_pair_c), same bits.
struct __align__(32) LinkRec {
    double rd_a, rd_b;
    uint64_t pk_a, pk_b;
};
static_assert(sizeof(LinkRec) == 32, "LinkRec must be one sector");
struct JoinBits {
    int ib_a, kb_a, ib_b, kb_b;  // id bits / key bits of TreeA's and TreeB's packed word
};
__device__ __forceinline__ LinkRec st_ld_linkrec(const LinkRec *p) {
    uint64_t w0, w1, w2, w3;
    asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3)
                 : "l"(p));
    return LinkRec{__longlong_as_double((long long)w0), __longlong_as_double((long long)w1), w2, w3};
}
// distance between the leaves of two links in one tree, from the joined halves alone
__device__ __forceinline__ double joined_query(const TreeView &tv, const SmemTables &sm, double rd1, uint64_t pk1,
                                               double rd2, uint64_t pk2, int ib, int kb) {
    const uint32_t idm = (1u << ib) - 1u, km = (1u << kb) - 1u;
    const int32_t a = int32_t(uint32_t(pk1) & idm), b = int32_t(uint32_t(pk2) & idm);
    if (a == b) return 0.0;
    const bool a_lo = a < b;
    const uint64_t pl = a_lo ? pk1 : pk2, ph = a_lo ? pk2 : pk1;
    const RecC l{a_lo ? rd1 : rd2, uint32_t(pl >> ib) & km, 0.0};
    const RecC h{a_lo ? rd2 : rd1, uint32_t(ph >> (ib + kb)) & km, 0.0};
    double d = 0.0;
    int32_t m = 0;
    st_pair_c<false>(tv, sm, PairQ{a_lo ? a : b, a_lo ? b : a, false}, l, h, true, false, d, m);
    return d;
}
// builds the joined records of `L` link rows (b, a)
__global__ void k_join_links(const TreeView ta, const TreeView tb, const int2 *__restrict__ rows, int64_t L,
                             JoinBits jb, LinkRec *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const int2 r = rows[i];
    const NodeRec16 a = ta.rec16[r.y], b = tb.rec16[r.x];
    LinkRec o;
    o.rd_a = a.rd;
    o.rd_b = b.rd;
    o.pk_a = uint64_t(uint32_t(r.y)) | (uint64_t(a.suf) << jb.ib_a) | (uint64_t(a.pre) << (jb.ib_a + jb.kb_a));
    o.pk_b = uint64_t(uint32_t(r.x)) | (uint64_t(b.suf) << jb.ib_b) | (uint64_t(b.pre) << (jb.ib_b + jb.kb_b));
    out[i] = o;
}

// layout mode of a tree as a template argument (0 wide, 1 compact, 3 wide records +
// 32-bit tables): the two-tree kernels are instantiated per pair of modes, so no
// run-time layout branches or dead table pointers cost registers
static inline int tree_mode(const st_tree *t) { return t->compact ? 1 : (t->compact_tables ? 3 : 0); }
#define ST_DISPATCH_MODES(ma, mb, CALL)                              \
    do {                                                             \
        const int _k = (ma) * 4 + (mb);                              \
        switch (_k) {                                                \
            case 0:  { CALL(0, 0); } break;                          \
            case 1:  { CALL(0, 1); } break;                          \
            case 3:  { CALL(0, 3); } break;                          \
            case 4:  { CALL(1, 0); } break;                          \
            case 5:  { CALL(1, 1); } break;                          \
            case 7:  { CALL(1, 3); } break;                          \
            case 12: { CALL(3, 0); } break;                          \
            case 13: { CALL(3, 1); } break;                          \
            default: { CALL(3, 3); } break;                          \
        }                                                            \
    } while (0)

// k = i(i-1)/2 + j, 0 <= j < i   (the reference's loop order, MuchTree.pyx:2919-2925)
__device__ __forceinline__ void tri_unrank(int64_t k, int64_t &i, int64_t &j) {
    int64_t r = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
    while (r * (r - 1) / 2 > k) --r;
    while ((r + 1) * r / 2 <= k) ++r;
    i = r;
    j = k - r * (r - 1) / 2;
}

// advance (i, j) by `step` positions of the enumeration (cheap for step <~ i)
__device__ __forceinline__ void tri_advance(int64_t &i, int64_t &j, int64_t step) {
    j += step;
    while (j >= i) {
        j -= i;
        ++i;
    }
}

// ids of one tree for link pairs [k0, k0+m): out[k-k0] = dist(col[j], col[i])
__global__ void __launch_bounds__(LT)
k_linked(const TreeView tv, const int32_t *__restrict__ col, int64_t k0, int64_t m,
         double *__restrict__ out, int64_t *__restrict__ ids_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar;
    const SmemTables sm = st_load_tables(tv, smem_raw, &tables_bar);
    // every warp owns a contiguous run of the enumeration: un-rank once per thread, then
    // step by 32 (lanes stay adjacent: coalesced link loads and result stores)
    const int64_t n_warps = int64_t(gridDim.x) * (LT / 32);
    const int64_t per_warp = ((m + n_warps - 1) / n_warps + 31) & ~int64_t(31);
    const int64_t wbeg = (int64_t(blockIdx.x) * (LT / 32) + (threadIdx.x >> 5)) * per_warp;
    const int64_t wend = wbeg + per_warp < m ? wbeg + per_warp : m;
    int64_t q = wbeg + (threadIdx.x & 31);
    int64_t i = 1, j = 0;
    if (q < wend) tri_unrank(k0 + q, i, j);
    for (; q < wend; q += 32, tri_advance(i, j, 32)) {
        int32_t a = __ldg(col + j), b = __ldg(col + i);
        st_st_stream_f64(out + q, linked_query(tv, sm, a, b));
        if (ids_out) {
            ids_out[2 * q] = a;
            ids_out[2 * q + 1] = b;
        }
    }
}

// ------------------------------------------------------------ link handle ---
// The link list lives on the device for as long as the caller keeps the handle: the
// per-call O(L) id conversion, the pageable H2D copies and their synchronisation are
// paid once (st_links_create), not by every sampler / moment / scan call.
struct st_links {
    const st_tree *ta = nullptr, *tb = nullptr;
    int device = 0;
    int64_t L = 0;
    int32_t *col_a = nullptr, *col_b = nullptr;  // linklist[:,1] (TreeA ids), linklist[:,0] (TreeB ids)
    int2 *rows = nullptr;                        // (b, a) per link, for the samplers
    LinkRec *joined = nullptr;                   // joined records of `rows` (NULL: layouts / field widths do not allow it)
    JoinBits jb{0, 0, 0, 0};
    // per-clade scans (st_links_clade_moments), built on first use per side (0: TreeB, 1: TreeA):
    // the links sorted (stably) by their id on that side, and first[v] = number of links with
    // side id < v -- a clade (an id interval) is then the run [first[lo], first[hi + 1])
    std::vector<int2> h_rows;
    mutable std::mutex scan_mu;
    mutable int2 *rows_sorted[2] = {nullptr, nullptr};
    mutable LinkRec *joined_sorted[2] = {nullptr, nullptr};  // joined records of rows_sorted[side]
    mutable int32_t *first[2] = {nullptr, nullptr};
};

static int check_pair_of_trees(const st_tree *ta, const st_tree *tb, const int64_t *linklist, int64_t L) {
    if (!ta || !tb || !linklist || L < 0) {
        st_set_error("linked: NULL tree / linklist or negative n_links");
        return ST_ERR_INVALID_ARG;
    }
    if (ta->device != tb->device) {
        st_set_error("linked: both trees must live on the same device (%d vs %d)", ta->device, tb->device);
        return ST_ERR_INVALID_ARG;
    }
    return ST_OK;
}

extern "C" void st_links_destroy(st_links *k) {
    if (!k) return;
    DeviceGuard g(k->device);
    cudaFree(k->col_a);
    cudaFree(k->col_b);
    cudaFree(k->rows);
    cudaFree(k->joined);
    for (int i = 0; i < 2; ++i) {
        cudaFree(k->joined_sorted[i]);
        cudaFree(k->rows_sorted[i]);
        cudaFree(k->first[i]);
    }
    delete k;
}

// field widths of one tree's packed word, or false when a joined record cannot hold the tree
static bool join_bits(const st_tree *t, int *ib, int *kb) {
    if (!t->compact) return false;
    *ib = std::max(1, st_ceil_log2_i64(t->n_nodes));
    *kb = std::min(31, (64 - *ib) / 2);
    if (*ib > 31) return false;
    // compact keys are depth << block_shift | offset in block, depth <= the tree's depth
    const uint64_t max_key = (uint64_t(uint32_t(t->depth)) << t->block_shift) | ((uint64_t(1) << t->block_shift) - 1);
    return max_key < (uint64_t(1) << *kb);
}
// SUCHTREE_B200_JOINED = 0 keeps the separate link rows + node records (tests, experiments); read per call
static bool joined_enabled() {
    const char *e = getenv("SUCHTREE_B200_JOINED");
    return !(e && e[0] == '0');
}
// joined records of L device rows into a fresh allocation (stream-ordered on the default stream, synchronised)
static int build_joined(const st_links *k, const int2 *d_rows, int64_t L, LinkRec **out) {
    *out = nullptr;
    if (L < 1) return ST_OK;
    LinkRec *d = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d), size_t(L) * sizeof(LinkRec)) != cudaSuccess) {
        cudaGetLastError();
        return ST_OK;  // not an error: the callers fall back to the separate rows
    }
    k_join_links<<<int((L + 255) / 256), 256>>>(k->ta->view, k->tb->view, d_rows, L, k->jb, d);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        st_set_error("st_links: building the joined link records failed");
        cudaFree(d);
        return ST_ERR_CUDA;
    }
    *out = d;
    return ST_OK;
}

extern "C" int st_links_create(const st_tree *ta, const st_tree *tb, const int64_t *linklist, int64_t L,
                               st_links **out) {
    if (!out) return ST_ERR_INVALID_ARG;
    *out = nullptr;
    int rc = check_pair_of_trees(ta, tb, linklist, L);
    if (rc != ST_OK) return rc;
    if (L >= (int64_t(1) << 31)) {
        st_set_error("st_links_create: n_links must be < 2^31");
        return ST_ERR_INVALID_ARG;
    }
    std::vector<int32_t> a(static_cast<size_t>(L)), b(static_cast<size_t>(L));
    std::vector<int2> rows(static_cast<size_t>(L));
    for (int64_t i = 0; i < L; ++i) {
        int64_t vb = linklist[2 * i], va = linklist[2 * i + 1];
        if (va < 0 || va >= ta->n_nodes) {
            st_set_bad_node(va);
            st_set_error("linklist row %lld: TreeA id %lld out of bounds", (long long)i, (long long)va);
            return ST_ERR_NODE_RANGE;
        }
        if (vb < 0 || vb >= tb->n_nodes) {
            st_set_bad_node(vb);
            st_set_error("linklist row %lld: TreeB id %lld out of bounds", (long long)i, (long long)vb);
            return ST_ERR_NODE_RANGE;
        }
        a[size_t(i)] = int32_t(va);
        b[size_t(i)] = int32_t(vb);
        rows[size_t(i)] = make_int2(int32_t(vb), int32_t(va));
    }
    DeviceGuard g(ta->device);
    st_links *k = new (std::nothrow) st_links();
    if (!k) return ST_ERR_NOMEM;
    k->ta = ta;
    k->tb = tb;
    k->device = ta->device;
    k->L = L;
    k->h_rows = rows;
    const size_t n = size_t(std::max<int64_t>(L, 1));
    if (cudaMalloc(reinterpret_cast<void **>(&k->col_a), n * 4) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&k->col_b), n * 4) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&k->rows), n * 8) != cudaSuccess ||
        cudaMemcpy(k->col_a, a.data(), size_t(L) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(k->col_b, b.data(), size_t(L) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(k->rows, rows.data(), size_t(L) * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        st_set_error("st_links_create: %s", cudaGetErrorString(cudaGetLastError()));
        st_links_destroy(k);
        return ST_ERR_CUDA;
    }
    if (join_bits(ta, &k->jb.ib_a, &k->jb.kb_a) && join_bits(tb, &k->jb.ib_b, &k->jb.kb_b)) {
        rc = build_joined(k, k->rows, L, &k->joined);
        if (rc != ST_OK) {
            st_links_destroy(k);
            return rc;
        }
    }
    *out = k;
    return ST_OK;
}

// entry points that take the raw link list build a handle for the call
struct TempLinks {
    st_links *k = nullptr;
    ~TempLinks() { st_links_destroy(k); }
};

// 6 moment sums {n, sx, sy, sxx, syy, sxy} at lane->d_scratch -> host, through the path's one
// collective when a communicator is given: ncclAllReduce(sum, fp64, 6) enqueued on the
// stream the kernels ran on (st_nccl.cu)
int st_nccl_allreduce_sum_f64(void *comm, double *d_buf, int count, cudaStream_t stream);
static int finish_moments(HostLane *lane, cudaStream_t s, void *nccl_comm, const char *who, double x0,
                          double y0, st_moments *out) {
    if (nccl_comm) {
        const int rc = st_nccl_allreduce_sum_f64(nccl_comm, lane->d_scratch, 6, s);
        if (rc != ST_OK) {
            cudaStreamSynchronize(s);
            return rc;
        }
    }
    cudaMemcpyAsync(lane->h_scratch, lane->d_scratch, 6 * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        st_set_error("%s: %s", who, cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    const double *h = lane->h_scratch;
    out->n = h[0];
    out->x0 = x0;
    out->y0 = y0;
    out->sx = h[1]; out->sy = h[2]; out->sxx = h[3]; out->syy = h[4]; out->sxy = h[5];
    return ST_OK;
}

// ------------------------------------------------------ linked_distances ----
extern "C" int st_links_linked_distances(const st_links *k, double *out_a, double *out_b, int64_t *ids_a,
                                         int64_t *ids_b) {
    if (!k) {
        st_set_error("st_links_linked_distances: NULL handle");
        return ST_ERR_INVALID_ARG;
    }
    const st_tree *ta = k->ta, *tb = k->tb;
    const int64_t L = k->L, total = L * (L - 1) / 2;
    if (total <= 0) return ST_OK;
    if (!out_a || !out_b) {
        st_set_error("st_linked_distances: NULL output");
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(ta->device);
    LaneGuard lg(ta->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    int rc = st_raise_smem(k_linked, ta->device, std::max(ta->query_smem_bytes, tb->query_smem_bytes));
    if (rc != ST_OK) return rc;

    const int64_t C = std::min<int64_t>(total, int64_t(1) << 22);
    double *d_out[2] = {nullptr, nullptr};
    int64_t *d_ids[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {lane->streams[0], lane->streams[1]};
    auto cleanup = [&]() {
        for (int i = 0; i < 2; ++i) {
            if (d_out[i]) cudaFreeAsync(d_out[i], st[i]);
            if (d_ids[i]) cudaFreeAsync(d_ids[i], st[i]);
        }
    };
    for (int i = 0; i < 2; ++i) {
        if (cudaMallocAsync(reinterpret_cast<void **>(&d_out[i]), size_t(C) * 8, st[i]) != cudaSuccess ||
            ((ids_a || ids_b) &&
             cudaMallocAsync(reinterpret_cast<void **>(&d_ids[i]), size_t(C) * 16, st[i]) != cudaSuccess)) {
            cudaGetLastError();
            cleanup();
            st_set_error("st_linked_distances: device allocation failed");
            return ST_ERR_NOMEM;
        }
    }
    int slot = 0;
    for (int tree = 0; tree < 2; ++tree) {
        const st_tree *t = tree ? tb : ta;
        const int32_t *col = tree ? k->col_b : k->col_a;
        double *out = tree ? out_b : out_a;
        int64_t *ids = tree ? ids_b : ids_a;
        for (int64_t k0 = 0; k0 < total; k0 += C, slot ^= 1) {
            const int64_t m = std::min(C, total - k0);
            int grid = int(std::min<int64_t>((m + LT - 1) / LT, int64_t(t->sm_count) * 8));
            k_linked<<<grid, LT, t->query_smem_bytes, st[slot]>>>(t->view, col, k0, m, d_out[slot],
                                                                  ids ? d_ids[slot] : nullptr);
            cudaMemcpyAsync(out + k0, d_out[slot], size_t(m) * 8, cudaMemcpyDeviceToHost, st[slot]);
            if (ids)
                cudaMemcpyAsync(ids + 2 * k0, d_ids[slot], size_t(m) * 16, cudaMemcpyDeviceToHost, st[slot]);
        }
    }
    cleanup();
    cudaError_t e0 = cudaStreamSynchronize(st[0]), e1 = cudaStreamSynchronize(st[1]);
    cudaError_t e2 = cudaGetLastError();
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) {
        st_set_error("st_linked_distances: %s",
                     cudaGetErrorString(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2)));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}

extern "C" int st_linked_distances(const st_tree *ta, const st_tree *tb, const int64_t *linklist,
                                   int64_t L, double *out_a, double *out_b, int64_t *ids_a,
                                   int64_t *ids_b) {
    TempLinks tl;
    int rc = st_links_create(ta, tb, linklist, L, &tl.k);
    if (rc != ST_OK) return rc;
    return st_links_linked_distances(tl.k, out_a, out_b, ids_a, ids_b);
}

// ------------------------------------------- xorshift64*: exact jump-ahead --
// The reference draws link indices from ONE sequential xorshift64* state
// (MuchTree.pyx:2946-2949).  The state update s ^= s>>12; s ^= s<<25; s ^= s>>27
// is linear over GF(2): s' = T s.  With the 64x64 bit matrices T^(2^k) a thread
// jumps straight to the state before ITS draws, so the device reproduces the
// reference's stream bit for bit, in parallel.
static const int XS_LEVELS = 48;
struct XsJump {
    uint64_t col[XS_LEVELS][64];  // col[k][b] = T^(2^k) applied to the unit vector e_b
};
static XsJump g_xs;
static bool g_xs_ready = false;
static std::mutex g_xs_mu;

static inline uint64_t xs_step(uint64_t s) {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s;
}
static inline uint64_t xs_apply(const uint64_t *col, uint64_t s) {
    uint64_t r = 0;
    for (int b = 0; b < 64; ++b)
        if ((s >> b) & 1) r ^= col[b];
    return r;
}
static const XsJump &xs_table() {
    std::lock_guard<std::mutex> lock(g_xs_mu);
    if (!g_xs_ready) {
        for (int b = 0; b < 64; ++b) g_xs.col[0][b] = xs_step(uint64_t(1) << b);
        for (int k = 1; k < XS_LEVELS; ++k)
            for (int b = 0; b < 64; ++b) g_xs.col[k][b] = xs_apply(g_xs.col[k - 1], g_xs.col[k - 1][b]);
        g_xs_ready = true;
    }
    return g_xs;
}
static uint64_t xs_jump_host(uint64_t s, uint64_t steps) {
    const XsJump &J = xs_table();
    for (int k = 0; k < XS_LEVELS && steps; ++k, steps >>= 1)
        if (steps & 1) s = xs_apply(J.col[k], s);
    return s;
}

__device__ __forceinline__ uint64_t xs_apply_dev(const uint64_t *__restrict__ col, uint64_t s) {
    uint64_t r = 0;
#pragma unroll 8
    for (int b = 0; b < 64; ++b) r ^= ((s >> b) & 1) ? __ldg(col + b) : 0ull;
    return r;
}

static const int XS_RUN = 8;  // consecutive samples per thread after one jump

// One pass over the cycle's samples for ONE tree.  which = 1: TreeA (linklist[:,1]),
// which = 0: TreeB (linklist[:,0]).  Sample q consumes draws 2q, 2q+1 of the stream.
__global__ void __launch_bounds__(LT)
k_sample_xs(const TreeView tv, const int2 *__restrict__ links, uint64_t n_links, int which,
            uint64_t seed, const uint64_t *__restrict__ jump, int64_t total, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar;
    const SmemTables sm = st_load_tables(tv, smem_raw, &tables_bar);
    const int64_t runs = (total + XS_RUN - 1) / XS_RUN;
    for (int64_t r = int64_t(blockIdx.x) * LT + threadIdx.x; r < runs; r += int64_t(gridDim.x) * LT) {
        const int64_t q0 = r * XS_RUN;
        uint64_t s = seed, steps = 2ull * uint64_t(q0);
        for (int k = 0; steps && k < XS_LEVELS; ++k, steps >>= 1)
            if (steps & 1) s = xs_apply_dev(jump + k * 64, s);
        for (int u = 0; u < XS_RUN && q0 + u < total; ++u) {
            s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
            int l1 = int((s * 2685821657736338717ull) % n_links);  // `cdef int l1` (MuchTree.pyx:3009)
            s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
            int l2 = int((s * 2685821657736338717ull) % n_links);
            int2 r1 = __ldg(links + l1), r2 = __ldg(links + l2);
            out[q0 + u] = which ? linked_query(tv, sm, r1.y, r2.y) : linked_query(tv, sm, r1.x, r2.x);
        }
    }
}

// Per-bucket sum d and sum d^2 (MuchTree.pyx:3045-3052) from the materialised distances, in a
// FIXED order: one CTA per (bucket, tree); thread t adds elements t, t+256, ... in order, then a
// fixed shuffle / shared-memory tree.  Bit-reproducible run to run (the reference's bucket
// statistics are): the `deviation < sigma` stopping test can never flip between runs.
__global__ void __launch_bounds__(256)
k_bucket_sums(const double *__restrict__ d, int32_t n_per_bucket, int32_t buckets,
              double *__restrict__ stats /* [tree][sum|sumsq][buckets] */) {
    const int b = blockIdx.x, tree = blockIdx.y;
    const double *x = d + (size_t(tree) * buckets + b) * n_per_bucket;
    double s = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n_per_bucket; i += 256) {
        const double v = x[i];
        s += v;
        s2 += v * v;
    }
    __shared__ double red[2][8];
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        stats[(size_t(tree) * 2 + threadIdx.x) * buckets + b] = t;
    }
}

extern "C" int st_links_sample_cycle(const st_links *k, uint64_t *seed, int32_t buckets, int32_t n,
                                     double *out_a, double *out_b, double *sums_a, double *sumsq_a,
                                     double *sums_b, double *sumsq_b) {
    if (!k) {
        st_set_error("st_links_sample_cycle: NULL handle");
        return ST_ERR_INVALID_ARG;
    }
    const st_tree *ta = k->ta, *tb = k->tb;
    const int64_t L = k->L;
    if (!seed || buckets < 1 || n < 1 || L < 1 || !out_a || !out_b || !sums_a || !sumsq_a || !sums_b ||
        !sumsq_b) {
        st_set_error("st_sample_linked_cycle: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    const int64_t total = int64_t(buckets) * n;
    if (2ull * uint64_t(total) >= (uint64_t(1) << XS_LEVELS)) {  // the jump table has XS_LEVELS levels
        st_set_error("st_sample_linked_cycle: 2*buckets*n must be < 2^%d", XS_LEVELS);
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(ta->device);
    LaneGuard lg(ta->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    int rc = st_raise_smem(k_sample_xs, ta->device, std::max(ta->query_smem_bytes, tb->query_smem_bytes));
    if (rc != ST_OK) return rc;
    const XsJump &J = xs_table();
    uint64_t *d_jump = nullptr;
    double *d_out = nullptr, *d_stats = nullptr;
    cudaStream_t s = lane->streams[0];
    auto cleanup = [&]() {  // stream-ordered scratch: recycled by the pool, no driver malloc per cycle
        if (d_jump) cudaFreeAsync(d_jump, s);
        if (d_out) cudaFreeAsync(d_out, s);
        if (d_stats) cudaFreeAsync(d_stats, s);
    };
    if (cudaMallocAsync(reinterpret_cast<void **>(&d_jump), sizeof(J.col), s) != cudaSuccess ||
        cudaMallocAsync(reinterpret_cast<void **>(&d_out), size_t(total) * 8 * 2, s) != cudaSuccess ||
        cudaMallocAsync(reinterpret_cast<void **>(&d_stats), size_t(buckets) * 8 * 4, s) != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        st_set_error("st_sample_linked_cycle: device allocation failed");
        return ST_ERR_NOMEM;
    }
    cudaMemcpyAsync(d_jump, J.col, sizeof(J.col), cudaMemcpyHostToDevice, s);
    const int64_t runs = (total + XS_RUN - 1) / XS_RUN;
    for (int tree = 0; tree < 2; ++tree) {
        const st_tree *t = tree ? tb : ta;
        int grid = int(std::min<int64_t>((runs + LT - 1) / LT, int64_t(t->sm_count) * 8));
        k_sample_xs<<<grid, LT, t->query_smem_bytes, s>>>(t->view, k->rows, uint64_t(L), tree ? 0 : 1, *seed,
                                                          d_jump, total, d_out + tree * total);
    }
    k_bucket_sums<<<dim3(unsigned(buckets), 2), 256, 0, s>>>(d_out, n, buckets, d_stats);
    std::vector<double> stats(size_t(buckets) * 4);
    // pageable result arrays (the Python shim's): one D2H of both vectors into the lane's page-locked
    // staging, then a parallel copy out -- a pageable cudaMemcpy of 2 x 2 MB per cycle is what the
    // cycle used to spend most of its time in
    const bool direct = st_is_pinned(out_a) && st_is_pinned(out_b);
    double *h_stage = nullptr;
    if (!direct && 2 * total <= ST_STAGE_PAIRS_MAX && st_lane_ensure_stage(lane, 2 * total, false, true) == ST_OK)
        h_stage = static_cast<double *>(lane->h_out[0]);
    if (h_stage) {
        cudaMemcpyAsync(h_stage, d_out, size_t(total) * 16, cudaMemcpyDeviceToHost, s);
    } else {
        cudaMemcpyAsync(out_a, d_out, size_t(total) * 8, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(out_b, d_out + total, size_t(total) * 8, cudaMemcpyDeviceToHost, s);
    }
    cudaMemcpyAsync(stats.data(), d_stats, stats.size() * 8, cudaMemcpyDeviceToHost, s);
    cleanup();
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        st_set_error("st_sample_linked_cycle: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    if (h_stage) {
        st_parallel_copy(out_a, h_stage, size_t(total) * 8);
        st_parallel_copy(out_b, h_stage + total, size_t(total) * 8);
    }
    for (int i = 0; i < buckets; ++i) {  // stats: [TreeA sum | TreeA sumsq | TreeB sum | TreeB sumsq][buckets]
        sums_a[i] += stats[i];
        sumsq_a[i] += stats[buckets + i];
        sums_b[i] += stats[2 * buckets + i];
        sumsq_b[i] += stats[3 * buckets + i];
    }
    *seed = xs_jump_host(*seed, 2ull * uint64_t(total));
    return ST_OK;
}

extern "C" int st_sample_linked_cycle(const st_tree *ta, const st_tree *tb, const int64_t *linklist,
                                      int64_t L, uint64_t *seed, int32_t buckets, int32_t n,
                                      double *out_a, double *out_b, double *sums_a, double *sumsq_a,
                                      double *sums_b, double *sumsq_b) {
    TempLinks tl;
    int rc = st_links_create(ta, tb, linklist, L, &tl.k);
    if (rc != ST_OK) return rc;
    return st_links_sample_cycle(tl.k, seed, buckets, n, out_a, out_b, sums_a, sumsq_a, sums_b, sumsq_b);
}

// ------------------------------------------------ Philox sampler + moments --
// Throughput path for the sampled two-tree correlation: nothing is
// materialised.  Per sample: 2 link rows + 2 x (3 index sectors); moments are
// kept in registers, reduced by warp shuffles, one partial per CTA.
struct Mom5 {
    double sx, sy, sxx, syy, sxy;
};
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// both trees' block tables live in shared memory (up to ~2 x 96 KB): one big CTA per SM
static const int MLT = 1024;

// J: joined link records (LinkRec, both trees compact) instead of link rows + node records
template <int MA, int MB, bool J = false>
__global__ void __launch_bounds__(MLT)
k_sample_moments(const TreeView ta, const TreeView tb, const void *__restrict__ links_v, JoinBits jb,
                 uint32_t n_links, uint64_t seed, int64_t first, int64_t n, double x0, double y0,
                 double *__restrict__ partials /* [grid][5] */) {
    const int2 *__restrict__ links = static_cast<const int2 *>(links_v);
    const LinkRec *__restrict__ joined = static_cast<const LinkRec *>(links_v);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // two table sets back to back (second one 16-byte aligned)
    __shared__ __align__(8) uint64_t tables_bar[2];
    const int offs = (st_table_bytes(ta.n_blocks, ta.st_levels, st_table_mode(ta)) + 15) & ~15;
    st_tables_issue(ta, smem_raw, &tables_bar[0]);  // both bulk copies in flight together
    st_tables_issue(tb, smem_raw + offs, &tables_bar[1]);
    __syncthreads();
    const SmemTables sa = st_tables_wait<MA>(ta, smem_raw, &tables_bar[0]);
    const SmemTables sb = st_tables_wait<MB>(tb, smem_raw + offs, &tables_bar[1]);
    __shared__ double red[5][MLT / 32];

    Mom5 m{0, 0, 0, 0, 0};
    // samples are processed two at a time: one Philox call = 4 words = 2 samples
    const int64_t c_begin = first >> 1, c_end = (first + n + 1) >> 1;
    for (int64_t c = c_begin + int64_t(blockIdx.x) * MLT + threadIdx.x; c < c_end;
         c += int64_t(gridDim.x) * MLT) {
        Philox4 r = st_philox4x32_10(uint64_t(c), seed);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t s = 2 * c + h;
            if (s < first || s >= first + n) continue;
            double x, y;
            if constexpr (J) {
                const LinkRec r1 = st_ld_linkrec(joined + st_bounded(w[2 * h], n_links));
                const LinkRec r2 = st_ld_linkrec(joined + st_bounded(w[2 * h + 1], n_links));
                x = joined_query(ta, sa, r1.rd_a, r1.pk_a, r2.rd_a, r2.pk_a, jb.ib_a, jb.kb_a) - x0;
                y = joined_query(tb, sb, r1.rd_b, r1.pk_b, r2.rd_b, r2.pk_b, jb.ib_b, jb.kb_b) - y0;
            } else {
                int2 l1 = __ldg(links + st_bounded(w[2 * h], n_links));
                int2 l2 = __ldg(links + st_bounded(w[2 * h + 1], n_links));
                x = linked_query<MA>(ta, sa, l1.y, l2.y) - x0;
                y = linked_query<MB>(tb, sb, l1.x, l2.x) - y0;
            }
            m.sx += x; m.sy += y;
            m.sxx += x * x; m.syy += y * y; m.sxy += x * y;
        }
    }
    double v[5] = {m.sx, m.sy, m.sxx, m.syy, m.sxy};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w2 = 0; w2 < MLT / 32; ++w2) s += red[threadIdx.x][w2];
        partials[size_t(blockIdx.x) * 5 + threadIdx.x] = s;
    }
}

// out[6] = {n, sx, sy, sxx, syy, sxy}: the payload of the moment all-reduce
__global__ void k_reduce_partials(int grid, const double *__restrict__ partials, double n,
                                  double *__restrict__ out) {
    // one warp per moment, fixed order -> deterministic
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= 5) return;
    double s = 0;
    for (int i = lane; i < grid; i += 32) s += partials[size_t(i) * 5 + k];
    s = warp_sum(s);
    if (lane == 0) out[k + 1] = s;
    if (threadIdx.x == 0) out[0] = n;
}

extern "C" int st_links_sample_moments(const st_links *k, uint64_t seed, int64_t first_sample,
                                       int64_t n_samples, double x0, double y0, void *nccl_comm,
                                       st_moments *out) {
    if (!k || !out || k->L < 1 || n_samples < 0 || first_sample < 0) {
        st_set_error("st_sample_moments: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    const st_tree *ta = k->ta, *tb = k->tb;
    out->n = double(n_samples);
    out->x0 = x0;
    out->y0 = y0;
    out->sx = out->sy = out->sxx = out->syy = out->sxy = 0.0;
    if (n_samples == 0 && !nccl_comm) return ST_OK;  // (with a communicator every rank must enter the collective)
    DeviceGuard g(ta->device);
    LaneGuard lg(ta->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    const int smem = ((ta->query_smem_bytes + 15) & ~15) + tb->query_smem_bytes;
    // 1024-thread CTAs with both trees' tables: one CTA per SM
    const int64_t calls = (n_samples + 1) / 2 + 1;
    int grid = int(std::min<int64_t>((calls + MLT - 1) / MLT, int64_t(ta->sm_count)));
    cudaStream_t s = lane->streams[0];
    double *d_part = nullptr;  // [grid][5] partials
    ST_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_part), size_t(grid) * 5 * 8, s));
    int rc_launch = ST_OK;
#define ST_LAUNCH_SAMPLE(MA, MB)                                                                        \
    rc_launch = st_raise_smem(k_sample_moments<MA, MB>, ta->device, smem);                              \
    if (rc_launch == ST_OK)                                                                             \
        k_sample_moments<MA, MB><<<grid, MLT, smem, s>>>(ta->view, tb->view, k->rows, k->jb, uint32_t(k->L), seed, \
                                                         first_sample, n_samples, x0, y0, d_part)
    if (k->joined && joined_enabled()) {
        rc_launch = st_raise_smem(k_sample_moments<1, 1, true>, ta->device, smem);
        if (rc_launch == ST_OK)
            k_sample_moments<1, 1, true><<<grid, MLT, smem, s>>>(ta->view, tb->view, k->joined, k->jb, uint32_t(k->L),
                                                                 seed, first_sample, n_samples, x0, y0, d_part);
    } else {
        ST_DISPATCH_MODES(tree_mode(ta), tree_mode(tb), ST_LAUNCH_SAMPLE);
    }
#undef ST_LAUNCH_SAMPLE
    if (rc_launch != ST_OK) {  // the kernel was not launched: no result to report
        cudaFreeAsync(d_part, s);
        return rc_launch;
    }
    k_reduce_partials<<<1, 160, 0, s>>>(grid, d_part, double(n_samples), lane->d_scratch);
    cudaFreeAsync(d_part, s);
    return finish_moments(lane, s, nccl_comm, "st_sample_moments", x0, y0, out);
}

extern "C" int st_sample_moments(const st_tree *ta, const st_tree *tb, const int64_t *linklist,
                                 int64_t L, uint64_t seed, int64_t first_sample, int64_t n_samples,
                                 double x0, double y0, st_moments *out) {
    TempLinks tl;
    int rc = st_links_create(ta, tb, linklist, L, &tl.k);
    if (rc != ST_OK) return rc;
    return st_links_sample_moments(tl.k, seed, first_sample, n_samples, x0, y0, nullptr, out);
}

// ------------------------------------- exhaustive link pairs, fused moments --
// linked_distances() followed by pearson() without the two distance vectors: the
// moments of (d_A, d_B) over link pairs k in [first, first + n) of the reference's
// enumeration (MuchTree.pyx:2919-2925), reduced in registers / shuffles / one partial
// per CTA.  This is the inner loop of the reference's per-clade correlation scan
// (docs/examples/SuchLinkedTree_examples.md:299-310).  Shardable by k-range.
template <int MA, int MB, bool J = false>
__global__ void __launch_bounds__(MLT)
k_linked_moments(const TreeView ta, const TreeView tb, const void *__restrict__ links_v, JoinBits jb, int64_t first,
                 int64_t n, double x0, double y0, double *__restrict__ partials /* [grid][5] */) {
    const int2 *__restrict__ links = static_cast<const int2 *>(links_v);
    const LinkRec *__restrict__ joined = static_cast<const LinkRec *>(links_v);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar[2];
    const int offs = (st_table_bytes(ta.n_blocks, ta.st_levels, st_table_mode(ta)) + 15) & ~15;
    st_tables_issue(ta, smem_raw, &tables_bar[0]);  // both bulk copies in flight together
    st_tables_issue(tb, smem_raw + offs, &tables_bar[1]);
    __syncthreads();
    const SmemTables sa = st_tables_wait<MA>(ta, smem_raw, &tables_bar[0]);
    const SmemTables sb = st_tables_wait<MB>(tb, smem_raw + offs, &tables_bar[1]);
    __shared__ double red[5][MLT / 32];

    Mom5 m{0, 0, 0, 0, 0};
    // every warp owns a contiguous run of the enumeration: un-rank once per thread, then
    // step by 32 (lanes stay adjacent: coalesced link loads)
    const int64_t n_warps = int64_t(gridDim.x) * (MLT / 32);
    const int64_t per_warp = ((n + n_warps - 1) / n_warps + 31) & ~int64_t(31);
    const int64_t wbeg = (int64_t(blockIdx.x) * (MLT / 32) + (threadIdx.x >> 5)) * per_warp;
    const int64_t wend = wbeg + per_warp < n ? wbeg + per_warp : n;
    int64_t q = wbeg + (threadIdx.x & 31);
    int64_t i = 1, j = 0;
    if (q < wend) tri_unrank(first + q, i, j);
    int64_t row = -1;  // the row (link i) whose records are cached below
    int2 l2 = make_int2(0, 0);
    RecFull ra{0.0, 0u, 0u}, rb{0.0, 0u, 0u};
    LinkRec ri{0.0, 0.0, 0ull, 0ull};
    for (; q < wend; q += 32, tri_advance(i, j, 32)) {
        double x, y;
        if constexpr (J) {
            if (i != row) {
                row = i;
                ri = st_ld_linkrec(joined + i);
            }
            const LinkRec rj = st_ld_linkrec(joined + j);  // lanes adjacent in j: contiguous records
            x = joined_query(ta, sa, ri.rd_a, ri.pk_a, rj.rd_a, rj.pk_a, jb.ib_a, jb.kb_a) - x0;
            y = joined_query(tb, sb, ri.rd_b, ri.pk_b, rj.rd_b, rj.pk_b, jb.ib_b, jb.kb_b) - y0;
        } else {
            if (i != row) {
                row = i;
                l2 = __ldg(links + i);
                if (MA == 1) ra = st_ld_rec_full(ta, l2.y);
                if (MB == 1) rb = st_ld_rec_full(tb, l2.x);
            }
            const int2 l1 = __ldg(links + j);
            x = linked_query_row<MA>(ta, sa, l2.y, ra, l1.y) - x0;
            y = linked_query_row<MB>(tb, sb, l2.x, rb, l1.x) - y0;
        }
        m.sx += x; m.sy += y;
        m.sxx += x * x; m.syy += y * y; m.sxy += x * y;
    }
    double v[5] = {m.sx, m.sy, m.sxx, m.syy, m.sxy};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w2 = 0; w2 < MLT / 32; ++w2) s += red[threadIdx.x][w2];
        partials[size_t(blockIdx.x) * 5 + threadIdx.x] = s;
    }
}

extern "C" int st_links_linked_moments(const st_links *k, int64_t first_pair, int64_t n_pairs, double x0,
                                       double y0, void *nccl_comm, st_moments *out) {
    if (!k || !out) {
        st_set_error("st_linked_moments: NULL handle or output");
        return ST_ERR_INVALID_ARG;
    }
    const st_tree *ta = k->ta, *tb = k->tb;
    const int64_t L = k->L, total = L * (L - 1) / 2;
    if (first_pair < 0 || n_pairs < 0 || first_pair + n_pairs > total) {
        st_set_error("st_linked_moments: bad arguments (pairs [%lld, %lld) of %lld)", (long long)first_pair,
                     (long long)(first_pair + n_pairs), (long long)total);
        return ST_ERR_INVALID_ARG;
    }
    out->n = double(n_pairs);
    out->x0 = x0;
    out->y0 = y0;
    out->sx = out->sy = out->sxx = out->syy = out->sxy = 0.0;
    if (n_pairs == 0 && !nccl_comm) return ST_OK;
    DeviceGuard g(ta->device);
    LaneGuard lg(ta->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    const int smem = ((ta->query_smem_bytes + 15) & ~15) + tb->query_smem_bytes;
    int grid = int(std::max<int64_t>(1, std::min<int64_t>((n_pairs + MLT - 1) / MLT, int64_t(ta->sm_count))));
    cudaStream_t s = lane->streams[0];
    double *d_part = nullptr;
    ST_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_part), size_t(grid) * 5 * 8, s));
    int rc_launch = ST_OK;
#define ST_LAUNCH_LINKED(MA, MB)                                                                 \
    rc_launch = st_raise_smem(k_linked_moments<MA, MB>, ta->device, smem);                       \
    if (rc_launch == ST_OK)                                                                      \
        k_linked_moments<MA, MB><<<grid, MLT, smem, s>>>(ta->view, tb->view, k->rows, k->jb, first_pair, \
                                                         n_pairs, x0, y0, d_part)
    if (k->joined && joined_enabled()) {
        rc_launch = st_raise_smem(k_linked_moments<1, 1, true>, ta->device, smem);
        if (rc_launch == ST_OK)
            k_linked_moments<1, 1, true><<<grid, MLT, smem, s>>>(ta->view, tb->view, k->joined, k->jb, first_pair,
                                                                 n_pairs, x0, y0, d_part);
    } else {
        ST_DISPATCH_MODES(tree_mode(ta), tree_mode(tb), ST_LAUNCH_LINKED);
    }
#undef ST_LAUNCH_LINKED
    if (rc_launch != ST_OK) {  // the kernel was not launched: no result to report
        cudaFreeAsync(d_part, s);
        return rc_launch;
    }
    k_reduce_partials<<<1, 160, 0, s>>>(grid, d_part, double(n_pairs), lane->d_scratch);
    cudaFreeAsync(d_part, s);
    return finish_moments(lane, s, nccl_comm, "st_linked_moments", x0, y0, out);
}

extern "C" int st_linked_moments(const st_tree *ta, const st_tree *tb, const int64_t *linklist,
                                 int64_t L, int64_t first_pair, int64_t n_pairs, double x0, double y0,
                                 st_moments *out) {
    TempLinks tl;
    int rc = st_links_create(ta, tb, linklist, L, &tl.k);
    if (rc != ST_OK) return rc;
    return st_links_linked_moments(tl.k, first_pair, n_pairs, x0, y0, nullptr, out);
}

// ----------------------------------------------- per-clade scan, one launch --
// The reference's co-phylogeny scan (docs/examples/SuchLinkedTree_examples.md:286-310) is a
// host loop over the internal nodes of one tree: subset_b(node) (MuchTree.pyx:2876-2886, which
// re-runs _build_linklist, :2845-2874), linked_distances() (:2900-2934), pearson().  A clade
// is a contiguous interval of in-order ids (SuchTree.get_leaves, :427-463), so with the links
// sorted by the id on the scanned side every subset is a contiguous RUN of links, and the whole
// scan is one flat list of work items (clade, pair range) over the same fused-moment inner loop
// as k_linked_moments.  One warp per item, items handed out by an atomic counter (any
// assignment gives the same result: each item owns its partial, folded per clade in item order).
struct CladeItem {
    int64_t first;  // first pair of the item in the clade's own (i, j<i) enumeration
    int32_t clade;  // index of the clade in the caller's arrays
    int32_t len;    // pairs in the item
};
static_assert(sizeof(CladeItem) == 16, "CladeItem is one 16-byte load");

// The PLAN runs on the device too (round 1 built it on the host: a counting sort of the links,
// run bounds, the eligible clades and the item list, per call): the sorted links and the prefix
// counts belong to the st_links handle (built once per side), and per call
//   k_clade_runs   clade -> run of links, eligibility, pair count; total pair count
//   k_clade_chunk  item size from the total (at most ~2^20 items, at least 4096 pairs each)
//   k_clade_count  items per clade  -> exclusive scan -> first item of every clade
//   k_clade_items  one warp per clade writes its items
// One 16-byte read-back (item count) sizes the partials; then shift | moments | fold as before.
struct CladeMeta {
    unsigned long long total_pairs;
    long long chunk;
    int32_t n_items, overflow;
};

__global__ void k_clade_runs(const int32_t *__restrict__ first, int64_t n_side, const int64_t *__restrict__ lo,
                             const int64_t *__restrict__ hi, int32_t n_clades, int64_t min_links,
                             int64_t max_links, int32_t *__restrict__ run_begin, int32_t *__restrict__ run_len,
                             int64_t *__restrict__ pairs, CladeMeta *meta) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (c < n_clades) {
        // links with lo <= id <= hi (bounds clamped to the tree)
        const int64_t l = lo[c] > 0 ? lo[c] : 0, h = hi[c] < n_side - 1 ? hi[c] : n_side - 1;
        const int32_t s = l <= h ? first[l] : 0;
        const int64_t n = l <= h ? int64_t(first[h + 1]) - s : 0;
        run_begin[c] = s;
        run_len[c] = int32_t(n);
        const bool elig = n >= min_links && n <= max_links;
        const int64_t p = elig ? n * (n - 1) / 2 : 0;
        pairs[c] = p;
        mine = (unsigned long long)p;
    }
    // integer sum: exact, order-independent
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&meta->total_pairs, mine);
}

__global__ void k_clade_chunk(CladeMeta *meta) {
    // items of at most `chunk` pairs (a multiple of 32), about 2^20 of them at most
    long long chunk = ((long long)(meta->total_pairs >> 20) + 31) & ~31ll;
    chunk = chunk < 4096 ? 4096 : (chunk > (1ll << 30) ? (1ll << 30) : chunk);
    meta->chunk = chunk;
}

__global__ void k_clade_count(const int64_t *__restrict__ pairs, int32_t n_clades, const CladeMeta *meta,
                              int32_t *__restrict__ items) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clades) return;
    const long long chunk = meta->chunk;
    items[c] = int32_t((pairs[c] + chunk - 1) / chunk);
}

// ---- exclusive scan of int32 (1024 elements per CTA, recursive over the CTA sums) ----
__global__ void __launch_bounds__(1024)
k_scan_block(const int32_t *__restrict__ in, int32_t n, int32_t *__restrict__ out, int32_t *__restrict__ sums) {
    __shared__ int32_t warp_tot[32];
    const int32_t i = blockIdx.x * 1024 + threadIdx.x;
    const int32_t v = i < n ? in[i] : 0;
    int32_t x = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        int32_t t = warp_tot[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        warp_tot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const int32_t before = w ? warp_tot[w - 1] : 0;
    if (i < n) out[i] = before + x - v;  // exclusive
    if (threadIdx.x == 1023 && sums) sums[blockIdx.x] = before + x;
}
__global__ void k_scan_add(int32_t *__restrict__ out, int32_t n, const int32_t *__restrict__ offs) {
    const int32_t i = blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += offs[blockIdx.x];
}
// scratch: room for ceil(n/1024) + ceil(n/1024^2) + ... + 1 ints (<= n/1000 + 8)
static void exclusive_scan_i32(const int32_t *in, int32_t n, int32_t *out, int32_t *scratch, cudaStream_t s) {
    const int32_t nb = (n + 1023) / 1024;
    k_scan_block<<<nb, 1024, 0, s>>>(in, n, out, nb > 1 ? scratch : nullptr);
    if (nb > 1) {
        exclusive_scan_i32(scratch, nb, scratch, scratch + nb, s);  // in place: every CTA reads before it writes
        k_scan_add<<<nb, 1024, 0, s>>>(out, n, scratch);
    }
}

// total item count (last exclusive prefix + last count); more than 2^31 items cannot be indexed
__global__ void k_clade_total(const int32_t *__restrict__ item_begin, const int32_t *__restrict__ items,
                              int32_t n_clades, CladeMeta *meta) {
    const long long t = (long long)item_begin[n_clades - 1] + items[n_clades - 1];
    meta->n_items = int32_t(t);
    meta->overflow = t > 0x7fffffffll || t < 0;
}

__global__ void k_clade_items(const int64_t *__restrict__ pairs, const int32_t *__restrict__ item_begin,
                              const int32_t *__restrict__ n_items_of, int32_t n_clades, const CladeMeta *meta,
                              CladeItem *__restrict__ items) {
    const int32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= n_clades) return;
    const int32_t cnt = n_items_of[c];
    if (cnt == 0) return;
    const long long chunk = meta->chunk;
    const int64_t p = pairs[c];
    CladeItem *dst = items + item_begin[c];
    for (int32_t k = lane; k < cnt; k += 32) {
        const int64_t f = int64_t(k) * chunk;
        dst[k] = CladeItem{f, c, int32_t(p - f < chunk ? p - f : chunk)};
    }
}

// x0, y0 of every computed clade: the distances of its first link pair (conditioning shift, as in
// SuchLinkedTrees.linked_pearson)
__global__ void k_clade_shift(const TreeView ta, const TreeView tb, const int2 *__restrict__ links,
                              const int32_t *__restrict__ run_begin, const int32_t *__restrict__ n_items_of,
                              int32_t n_clades, double2 *__restrict__ shift) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clades) return;
    if (n_items_of[c] == 0) {
        shift[c] = make_double2(0.0, 0.0);
        return;
    }
    const int2 l1 = __ldg(links + run_begin[c]), l2 = __ldg(links + run_begin[c] + 1);
    shift[c] = make_double2(linked_query<2>(ta, st_global_tables(ta), l1.y, l2.y),
                            linked_query<2>(tb, st_global_tables(tb), l1.x, l2.x));
}

template <int MA, int MB, bool J = false>
__global__ void __launch_bounds__(MLT)
k_clade_moments(const TreeView ta, const TreeView tb, const void *__restrict__ links_v, JoinBits jb,
                const int32_t *__restrict__ run_begin, const double2 *__restrict__ shift,
                const CladeItem *__restrict__ items, int32_t n_items, int32_t *__restrict__ next_item,
                double *__restrict__ partials /* [n_items][5] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar[2];
    const int offs = (st_table_bytes(ta.n_blocks, ta.st_levels, st_table_mode(ta)) + 15) & ~15;
    st_tables_issue(ta, smem_raw, &tables_bar[0]);
    st_tables_issue(tb, smem_raw + offs, &tables_bar[1]);
    __syncthreads();
    const SmemTables sa = st_tables_wait<MA>(ta, smem_raw, &tables_bar[0]);
    const SmemTables sb = st_tables_wait<MB>(tb, smem_raw + offs, &tables_bar[1]);
    const int lane = threadIdx.x & 31;
    for (;;) {
        int32_t it = 0;
        if (lane == 0) it = atomicAdd(next_item, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(items + it));
        const int64_t first = (int64_t)((uint64_t(uint32_t(raw.y)) << 32) | uint32_t(raw.x));
        const int32_t clade = raw.z, len = raw.w;
        const int32_t rb0 = __ldg(run_begin + clade);
        const int2 *__restrict__ run = static_cast<const int2 *>(links_v) + rb0;
        const LinkRec *__restrict__ jrun = static_cast<const LinkRec *>(links_v) + rb0;
        const double2 sh = __ldg(shift + clade);
        Mom5 m{0, 0, 0, 0, 0};
        int32_t i = 1, j = 0;  // link indices inside the run (n_links < 2^31)
        if (lane < len) {
            int64_t i64, j64;
            tri_unrank(first + lane, i64, j64);
            i = int32_t(i64);
            j = int32_t(j64);
        }
        int32_t row = -1;  // the row (link i of the run) whose records are cached below
        int2 l2 = make_int2(0, 0);
        RecFull ra{0.0, 0u, 0u}, rb{0.0, 0u, 0u};
        LinkRec ri{0.0, 0.0, 0ull, 0ull};
        for (int32_t q = lane; q < len; q += 32) {
            double x, y;
            if constexpr (J) {
                if (i != row) {
                    row = i;
                    ri = st_ld_linkrec(jrun + i);
                }
                const LinkRec rj = st_ld_linkrec(jrun + j);  // lanes adjacent in j: contiguous records
                for (j += 32; j >= i; ++i) j -= i;  // 32 positions on in the (i, j<i) enumeration
                x = joined_query(ta, sa, ri.rd_a, ri.pk_a, rj.rd_a, rj.pk_a, jb.ib_a, jb.kb_a) - sh.x;
                y = joined_query(tb, sb, ri.rd_b, ri.pk_b, rj.rd_b, rj.pk_b, jb.ib_b, jb.kb_b) - sh.y;
            } else {
                if (i != row) {
                    row = i;
                    l2 = __ldg(run + i);
                    if (MA == 1) ra = st_ld_rec_full(ta, l2.y);
                    if (MB == 1) rb = st_ld_rec_full(tb, l2.x);
                }
                const int2 l1 = __ldg(run + j);
                for (j += 32; j >= i; ++i) j -= i;  // 32 positions on in the (i, j<i) enumeration
                x = linked_query_row<MA>(ta, sa, l2.y, ra, l1.y) - sh.x;
                y = linked_query_row<MB>(tb, sb, l2.x, rb, l1.x) - sh.y;
            }
            m.sx += x; m.sy += y;
            m.sxx += x * x; m.syy += y * y; m.sxy += x * y;
        }
        double v[5] = {m.sx, m.sy, m.sxx, m.syy, m.sxy};
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = warp_sum(v[k]);
        if (lane < 5) {
            // lane k keeps moment k (all lanes hold every sum after the xor-shuffles)
            double mine = v[0];
#pragma unroll
            for (int k = 1; k < 5; ++k) mine = lane == k ? v[k] : mine;
            partials[size_t(it) * 5 + lane] = mine;
        }
    }
}

// one warp per clade folds the partials of its items in a fixed order and writes the clade's
// st_moments row {n, x0, y0, sx, sy, sxx, syy, sxy} (all zero when the clade was not computed)
__global__ void k_clade_fold(const double *__restrict__ partials, const int32_t *__restrict__ item_begin,
                             const int32_t *__restrict__ n_items_of, const int64_t *__restrict__ pairs,
                             const double2 *__restrict__ shift, const int32_t *__restrict__ run_len,
                             int32_t n_clades, double *__restrict__ out /* [n_clades][8] */,
                             int64_t *__restrict__ n_links_out) {
    const int32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= n_clades) return;
    const int32_t b = item_begin[c], e = b + n_items_of[c];
    double v[5] = {0, 0, 0, 0, 0};
    for (int32_t it = b + lane; it < e; it += 32)
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] += partials[size_t(it) * 5 + k];
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
        double *o = out + size_t(c) * 8;
        const bool done = e > b;
        o[0] = done ? double(pairs[c]) : 0.0;
        o[1] = done ? shift[c].x : 0.0;
        o[2] = done ? shift[c].y : 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) o[3 + k] = done ? v[k] : 0.0;
        if (n_links_out) n_links_out[c] = run_len[c];
    }
}

// links sorted by the id on `side` + prefix counts, on the device, once per handle and side
static int ensure_scan_index(const st_links *k, int side) {
    std::lock_guard<std::mutex> lock(k->scan_mu);
    if (k->rows_sorted[side]) return ST_OK;
    const int64_t L = k->L;
    const int64_t n_side = side == 0 ? k->tb->n_nodes : k->ta->n_nodes;
    // counting sort over the node ids (stable), O(links + nodes)
    std::vector<int32_t> first(static_cast<size_t>(n_side) + 1, 0);
    auto key = [&](int64_t i) { return side == 0 ? k->h_rows[size_t(i)].x : k->h_rows[size_t(i)].y; };
    for (int64_t i = 0; i < L; ++i) ++first[size_t(key(i)) + 1];
    for (int64_t v = 0; v < n_side; ++v) first[size_t(v) + 1] += first[size_t(v)];
    std::vector<int2> rows(static_cast<size_t>(std::max<int64_t>(L, 1)));
    {
        std::vector<int32_t> next(first.begin(), first.end() - 1);
        for (int64_t i = 0; i < L; ++i) rows[size_t(next[size_t(key(i))]++)] = k->h_rows[size_t(i)];
    }
    int2 *d_rows = nullptr;
    int32_t *d_first = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d_rows), rows.size() * 8) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&d_first), first.size() * 4) != cudaSuccess ||
        cudaMemcpy(d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        st_set_error("st_clade_moments: building the scan index failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(d_rows);
        cudaFree(d_first);
        return ST_ERR_CUDA;
    }
    if (k->joined) {  // the same records in the sorted order (a clade's links are then contiguous records)
        const int rc = build_joined(k, d_rows, L, &k->joined_sorted[side]);
        if (rc != ST_OK) {
            cudaFree(d_rows);
            cudaFree(d_first);
            return rc;
        }
    }
    k->first[side] = d_first;
    k->rows_sorted[side] = d_rows;
    return ST_OK;
}

extern "C" int st_links_clade_moments(const st_links *k, int side, const int64_t *clade_lo,
                                      const int64_t *clade_hi, int64_t n_clades, int64_t min_links,
                                      int64_t max_links, st_moments *out, int64_t *n_links_out) {
    if (!k || (side != 0 && side != 1) || n_clades < 0 || (n_clades > 0 && (!clade_lo || !clade_hi || !out)) ||
        n_clades >= (int64_t(1) << 31)) {
        st_set_error("st_clade_moments: bad arguments (side must be 0 or 1, clade arrays and out non-NULL)");
        return ST_ERR_INVALID_ARG;
    }
    if (n_clades == 0) return ST_OK;
    if (min_links < 2) min_links = 2;  // a pair needs two links
    if (max_links < 0) max_links = INT64_MAX;
    const st_tree *ta = k->ta, *tb = k->tb;
    DeviceGuard g(ta->device);
    int rc = ensure_scan_index(k, side);
    if (rc != ST_OK) return rc;
    LaneGuard lg(ta->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    cudaStream_t s = lane->streams[0];
    const int64_t n_side = side == 0 ? tb->n_nodes : ta->n_nodes;
    const int32_t nc = int32_t(n_clades);

    // one stream-ordered scratch block for the plan
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t scan_ints = size_t(nc) / 1000 + 64;
    const size_t o_lo = 0, o_hi = o_lo + al(size_t(nc) * 8), o_pairs = o_hi + al(size_t(nc) * 8),
                 o_run = o_pairs + al(size_t(nc) * 8), o_len = o_run + al(size_t(nc) * 4),
                 o_cnt = o_len + al(size_t(nc) * 4), o_ibeg = o_cnt + al(size_t(nc) * 4),
                 o_scan = o_ibeg + al(size_t(nc) * 4), o_shift = o_scan + al(scan_ints * 4),
                 o_out = o_shift + al(size_t(nc) * 16), o_nl = o_out + al(size_t(nc) * 64),
                 o_meta = o_nl + al(size_t(nc) * 8), o_next = o_meta + 256, bytes = o_next + 256;
    unsigned char *d = nullptr;
    if (cudaMallocAsync(reinterpret_cast<void **>(&d), bytes, s) != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_clade_moments: device allocation of %zu bytes failed", bytes);
        return ST_ERR_NOMEM;
    }
    int64_t *d_lo = reinterpret_cast<int64_t *>(d + o_lo), *d_hi = reinterpret_cast<int64_t *>(d + o_hi);
    int64_t *d_pairs = reinterpret_cast<int64_t *>(d + o_pairs);
    int32_t *d_run = reinterpret_cast<int32_t *>(d + o_run), *d_len = reinterpret_cast<int32_t *>(d + o_len);
    int32_t *d_cnt = reinterpret_cast<int32_t *>(d + o_cnt), *d_ibeg = reinterpret_cast<int32_t *>(d + o_ibeg);
    int32_t *d_scan = reinterpret_cast<int32_t *>(d + o_scan);
    double2 *d_shift = reinterpret_cast<double2 *>(d + o_shift);
    double *d_out = reinterpret_cast<double *>(d + o_out);
    int64_t *d_nl = reinterpret_cast<int64_t *>(d + o_nl);
    CladeMeta *d_meta = reinterpret_cast<CladeMeta *>(d + o_meta);
    int32_t *d_next = reinterpret_cast<int32_t *>(d + o_next);
    unsigned char *d2 = nullptr;  // items + partials, sized after the plan
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    auto fail = [&](int code) {
        if (d2) cudaFreeAsync(d2, s);
        cudaFreeAsync(d, s);
        cudaStreamSynchronize(s);
        return code;
    };
    step(cudaMemcpyAsync(d_lo, clade_lo, size_t(nc) * 8, cudaMemcpyHostToDevice, s));
    step(cudaMemcpyAsync(d_hi, clade_hi, size_t(nc) * 8, cudaMemcpyHostToDevice, s));
    step(cudaMemsetAsync(d_meta, 0, sizeof(CladeMeta), s));
    step(cudaMemsetAsync(d_next, 0, 4, s));
    const int T = 256, gb = (nc + T - 1) / T;
    k_clade_runs<<<gb, T, 0, s>>>(k->first[side], n_side, d_lo, d_hi, nc, min_links, max_links, d_run, d_len,
                                  d_pairs, d_meta);
    k_clade_chunk<<<1, 1, 0, s>>>(d_meta);
    k_clade_count<<<gb, T, 0, s>>>(d_pairs, nc, d_meta, d_cnt);
    exclusive_scan_i32(d_cnt, nc, d_ibeg, d_scan, s);
    k_clade_total<<<1, 1, 0, s>>>(d_ibeg, d_cnt, nc, d_meta);
    CladeMeta *h_meta = reinterpret_cast<CladeMeta *>(lane->h_scratch);  // pinned, 64 doubles
    step(cudaMemcpyAsync(h_meta, d_meta, sizeof(CladeMeta), cudaMemcpyDeviceToHost, s));
    step(cudaStreamSynchronize(s));
    step(cudaGetLastError());
    if (e != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_clade_moments: %s", cudaGetErrorString(e));
        return fail(ST_ERR_CUDA);
    }
    if (h_meta->overflow || double(h_meta->total_pairs) >= 9.0e18) {
        st_set_error("st_clade_moments: too much work for one call (%.3g link pairs)", double(h_meta->total_pairs));
        return fail(ST_ERR_INVALID_ARG);
    }
    const int32_t n_items = h_meta->n_items;
    CladeItem *d_items = nullptr;
    double *d_part = nullptr;
    if (n_items > 0) {
        const size_t b_items = al(size_t(n_items) * 16);
        if (cudaMallocAsync(reinterpret_cast<void **>(&d2), b_items + size_t(n_items) * 40, s) != cudaSuccess) {
            cudaGetLastError();
            st_set_error("st_clade_moments: device allocation for %d work items failed", n_items);
            return fail(ST_ERR_NOMEM);
        }
        d_items = reinterpret_cast<CladeItem *>(d2);
        d_part = reinterpret_cast<double *>(d2 + b_items);
        const int2 *rows = k->rows_sorted[side];
        k_clade_items<<<int((int64_t(nc) * 32 + T - 1) / T), T, 0, s>>>(d_pairs, d_ibeg, d_cnt, nc, d_meta, d_items);
        k_clade_shift<<<gb, T, 0, s>>>(ta->view, tb->view, rows, d_run, d_cnt, nc, d_shift);
        const int smem = ((ta->query_smem_bytes + 15) & ~15) + tb->query_smem_bytes;
        const int grid = int(std::min<int64_t>((int64_t(n_items) + MLT / 32 - 1) / (MLT / 32), int64_t(ta->sm_count)));
        int rc2 = ST_OK;
#define ST_LAUNCH_CLADE(MA, MB)                                                                          \
    rc2 = st_raise_smem(k_clade_moments<MA, MB>, ta->device, smem);                                      \
    if (rc2 == ST_OK)                                                                                    \
        k_clade_moments<MA, MB><<<grid, MLT, smem, s>>>(ta->view, tb->view, rows, k->jb, d_run, d_shift, d_items, \
                                                        n_items, d_next, d_part)
        if (k->joined_sorted[side] && joined_enabled()) {
            rc2 = st_raise_smem(k_clade_moments<1, 1, true>, ta->device, smem);
            if (rc2 == ST_OK)
                k_clade_moments<1, 1, true><<<grid, MLT, smem, s>>>(ta->view, tb->view, k->joined_sorted[side], k->jb,
                                                                    d_run, d_shift, d_items, n_items, d_next, d_part);
        } else {
            ST_DISPATCH_MODES(tree_mode(ta), tree_mode(tb), ST_LAUNCH_CLADE);
        }
#undef ST_LAUNCH_CLADE
        if (rc2 != ST_OK) return fail(rc2);
    }
    k_clade_fold<<<int((int64_t(nc) * 32 + T - 1) / T), T, 0, s>>>(d_part, d_ibeg, d_cnt, d_pairs, d_shift, d_len, nc,
                                                                 d_out, n_links_out ? d_nl : nullptr);
    step(cudaGetLastError());
    // rows are laid out as st_moments: straight into the caller's arrays
    static_assert(sizeof(st_moments) == 64, "st_moments is 8 doubles");
    step(cudaMemcpyAsync(out, d_out, size_t(nc) * 64, cudaMemcpyDeviceToHost, s));
    if (n_links_out) step(cudaMemcpyAsync(n_links_out, d_nl, size_t(nc) * 8, cudaMemcpyDeviceToHost, s));
    if (d2) cudaFreeAsync(d2, s);
    cudaFreeAsync(d, s);
    step(cudaStreamSynchronize(s));
    if (e != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_clade_moments: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}

extern "C" int st_clade_moments(const st_tree *ta, const st_tree *tb, const int64_t *linklist, int64_t L,
                                int side, const int64_t *clade_lo, const int64_t *clade_hi,
                                int64_t n_clades, int64_t min_links, int64_t max_links, st_moments *out,
                                int64_t *n_links_out) {
    TempLinks tl;
    int rc = st_links_create(ta, tb, linklist, L, &tl.k);
    if (rc != ST_OK) return rc;
    return st_links_clade_moments(tl.k, side, clade_lo, clade_hi, n_clades, min_links, max_links, out, n_links_out);
}

extern "C" double st_moments_pearson(const st_moments *m) {
    if (!m || m->n <= 0) return 0.0;
    // centred second moments from shifted sums; the shift cancels exactly in exact arithmetic
    const double n = m->n;
    const double cxx = m->sxx - m->sx * m->sx / n;
    const double cyy = m->syy - m->sy * m->sy / n;
    const double cxy = m->sxy - m->sx * m->sy / n;
    return cxy / std::sqrt(cxx * cyy + 1.0e-20);  // guard of MuchTree.pyx:79
}
