// Tree handle: host-side validation of the reference's Node arrays, then the
// device-side build of everything the query kernels read.
//
// Replaces `cdef struct Node` / `Node* data` (MuchTree.pyx:55-60, 114, 160) and
// the depth pass of SuchTree.__init__ (MuchTree.pyx:218-225).
//
// Device layout (all buffers read-only after the build):
//   rec[n]        32 B/node: root distance (double-double), suffix/prefix argmin
//                 of depth within the node's RMQ block           -> 1 sector/lookup
//   depth[n]      int32 node depth (root 0); only the same-block slow path scans it
//   stk[K][nb]    sparse table of packed (depth,id) keys over blocks \  copied to shared
//   brd[nb]       root distance of each block's minimum node         /  memory by queries
//   mst[Km][nm]   packed-key sparse table over micro blocks (same-block slow path)
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>

#include "st_device.cuh"
#include "st_hostctx.cuh"

// ------------------------------------------------------------ error state ---
static thread_local char g_err[512] = "";
static thread_local int64_t g_bad_node = 0;

void st_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void st_set_bad_node(int64_t id) { g_bad_node = id; }

extern "C" const char *st_last_error(void) { return g_err; }
extern "C" int64_t st_bad_node(void) { return g_bad_node; }
extern "C" int st_version(void) { return 200; }
#ifndef ST_BUILD_ID
#define ST_BUILD_ID "unknown"
#endif
extern "C" const char *st_build_id(void) { return ST_BUILD_ID; }
extern "C" int st_device_count(int *count) {
    if (!count) return ST_ERR_INVALID_ARG;
    *count = 0;
    ST_CUDA(cudaGetDeviceCount(count));
    return ST_OK;
}

// ------------------------------------------------------------ build kernels -
// Pointer jumping: after round r every node knows its 2^r-th ancestor and the
// (depth, root-distance) contribution of the 2^r edges below it.  ceil(log2(max
// depth))+1 rounds, each a flat pass over n nodes -- the 10^6-deep caterpillar
// costs 21 rounds, same as a balanced tree of the same size.
__global__ void k_jump_init(int32_t n, const int32_t *__restrict__ parent,
                            const float *__restrict__ edge, int32_t *__restrict__ anc,
                            int32_t *__restrict__ dep, double *__restrict__ rhi,
                            double *__restrict__ rlo) {
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int32_t p = parent[v];
    anc[v] = p;
    dep[v] = p >= 0 ? 1 : 0;
    rhi[v] = p >= 0 ? double(edge[v]) : 0.0;  // root carries the reference's -1 sentinel: not an edge
    rlo[v] = 0.0;
}

__global__ void k_jump_round(int32_t n, const int32_t *__restrict__ anc_in,
                             const int32_t *__restrict__ dep_in, const double *__restrict__ rhi_in,
                             const double *__restrict__ rlo_in, int32_t *__restrict__ anc_out,
                             int32_t *__restrict__ dep_out, double *__restrict__ rhi_out,
                             double *__restrict__ rlo_out, int *__restrict__ live) {
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int32_t a = anc_in[v];
    int32_t d = dep_in[v];
    dd r{rhi_in[v], rlo_in[v]};
    int32_t a2 = -1;
    if (a >= 0) {
        d += dep_in[a];
        r = dd_add(r, dd{rhi_in[a], rlo_in[a]});
        a2 = anc_in[a];
        if (a2 >= 0) *live = 1;
    }
    anc_out[v] = a2;
    dep_out[v] = d;
    rhi_out[v] = r.hi;
    rlo_out[v] = r.lo;
}

__global__ void k_max_depth(int32_t n, const int32_t *__restrict__ dep, int *__restrict__ out) {
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    int d = v < n ? dep[v] : 0;
    for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(out, d);
}

// One CTA per RMQ block: prefix / suffix argmin of depth inside the block, node
// records, and the block minimum.  Thread t owns a contiguous chunk of the block.
__global__ void k_block_records(int32_t n, int block_shift, const int32_t *__restrict__ dep,
                                const double *__restrict__ rhi, const double *__restrict__ rlo,
                                NodeRec *__restrict__ rec, uint64_t *__restrict__ blockmin,
                                double2 *__restrict__ brd) {
    __shared__ uint64_t cmin[256];
    __shared__ uint64_t cpre[256];  // min over chunks [0, t)
    __shared__ uint64_t csuf[256];  // min over chunks (t, T)
    const int32_t B = 1 << block_shift;
    const int32_t base = blockIdx.x << block_shift;
    const int T = blockDim.x;  // <= 256, divides B or equals B
    const int32_t chunk = B / T;
    const int32_t s = base + threadIdx.x * chunk;
    const int32_t e = min(s + chunk, n);  // exclusive
    uint64_t m = ~0ull;
    for (int32_t i = s; i < e; ++i) m = st_min64(m, st_key(dep[i], i));
    cmin[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = ~0ull;
        for (int t = 0; t < T; ++t) {
            cpre[t] = run;
            run = st_min64(run, cmin[t]);
        }
        blockmin[blockIdx.x] = run;
        const int32_t arg = st_key_id(run);
        brd[blockIdx.x] = make_double2(rhi[arg], rlo[arg]);
        run = ~0ull;
        for (int t = T - 1; t >= 0; --t) {
            csuf[t] = run;
            run = st_min64(run, cmin[t]);
        }
    }
    __syncthreads();
    uint64_t run = cpre[threadIdx.x];
    for (int32_t i = s; i < e; ++i) {
        run = st_min64(run, st_key(dep[i], i));
        rec[i].pre = run;
    }
    run = csuf[threadIdx.x];
    for (int32_t i = e - 1; i >= s; --i) {
        run = st_min64(run, st_key(dep[i], i));
        rec[i].suf = run;
        rec[i].rd_hi = rhi[i];
        rec[i].rd_lo = rlo[i];
    }
}

// Sparse table of keys over the block minima, one CTA (<= 4096 blocks).
// stk[k][i] = min key over blocks [i, min(i + 2^k, nb)); level 0 (the block minima)
// is written by k_block_records.
__global__ void k_block_sparse_table(int nb, int levels, uint64_t *stk) {
    for (int k = 1; k < levels; ++k) {
        const uint64_t *prev = stk + size_t(k - 1) * nb;
        uint64_t *cur = stk + size_t(k) * nb;
        int half = 1 << (k - 1);
        for (int i = threadIdx.x; i < nb; i += blockDim.x)
            cur[i] = st_min64(prev[i], prev[min(i + half, nb - 1)]);
        __syncthreads();  // global writes by this CTA are visible to it after the barrier
    }
}

__global__ void k_micro_min(int32_t n, int micro_shift, int32_t n_micro,
                            const int32_t *__restrict__ dep, uint64_t *__restrict__ mst0) {
    int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_micro) return;
    int32_t s = j << micro_shift, e = min(s + (1 << micro_shift), n);
    uint64_t m = ~0ull;
    for (int32_t i = s; i < e; ++i) m = st_min64(m, st_key(dep[i], i));
    mst0[j] = m;
}

__global__ void k_micro_level(int32_t n_micro, int half, const uint64_t *__restrict__ prev,
                              uint64_t *__restrict__ cur) {
    int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_micro) return;
    cur[j] = st_min64(prev[j], prev[min(j + half, n_micro - 1)]);
}

// ------------------------------------------------------------ compaction ----
__global__ void k_any_nonzero(int32_t n, const double *__restrict__ v, int *__restrict__ flag) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && v[i] != 0.0) *flag = 1;
}

__global__ void k_compact_records(int32_t n, int bs, const NodeRec *__restrict__ rec,
                                  NodeRec16 *__restrict__ rec16) {
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const NodeRec r = rec[v];
    const uint32_t mask = (1u << bs) - 1u;
    NodeRec16 c;
    c.rd = r.rd_hi;
    c.suf = (uint32_t(r.suf >> 32) << bs) | (uint32_t(r.suf) & mask);
    c.pre = (uint32_t(r.pre >> 32) << bs) | (uint32_t(r.pre) & mask);
    rec16[v] = c;
}

__global__ void k_compact_tables(int nb, int levels, int bs, int ts, const uint64_t *__restrict__ stk,
                                 const double2 *__restrict__ brd, uint32_t *__restrict__ stk32,
                                 double *__restrict__ brd8, int32_t *__restrict__ bid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < levels * nb) {
        const uint64_t k = stk[i];
        stk32[i] = (uint32_t(k >> 32) << ts) | (uint32_t(k) >> bs);
    }
    if (i < nb) {
        brd8[i] = brd[i].x;
        bid[i] = st_key_id(stk[i]);
    }
}

// Which pair kernel suits this tree?  The paired-record kernel (st_ld_rec_c<true>) saves the
// third, dependent gather whenever the MRCA is the sector neighbour of an endpoint, and costs a
// few per cent when it is not.  Probe: N random leaf pairs; count those whose MRCA is NOT a block
// minimum (the plain kernel would gather rd[mrca]) and, of those, the ones a sector neighbour
// answers.  counts[0] = pairs needing the gather, counts[1] = answered by a neighbour.
__global__ void k_probe_paired(const TreeView tv, uint32_t n_leaves, int32_t n_pairs, unsigned int *counts) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const Philox4 r = st_philox4x32_10(uint64_t(i), 0x5eedull);
    const int32_t a = 2 * int32_t(st_bounded(r.x, n_leaves)), b = 2 * int32_t(st_bounded(r.y, n_leaves));
    if (a == b) return;
    const int32_t lo = min(a, b), hi = max(a, b);
    const RecRaw l = st_ld_rec<1>(tv, lo), h = st_ld_rec<1>(tv, hi);
    bool ft = false;
    const uint64_t k = st_rmq<1>(tv, st_global_tables(tv), lo, hi, l.suf, h.pre, &ft);
    if (ft) return;
    const int32_t id = st_key_id(k);
    if (id == lo || id == hi) return;
    atomicAdd(counts, 1u);
    const bool lo_upper = (reinterpret_cast<uintptr_t>(tv.rec16 + lo) & 16) != 0;
    const bool hi_upper = (reinterpret_cast<uintptr_t>(tv.rec16 + hi) & 16) != 0;
    if (id == lo + (lo_upper ? -1 : 1) || id == hi + (hi_upper ? -1 : 1)) atomicAdd(counts + 1, 1u);
}

// ------------------------------------------------------------ validation ----
// Host check that (parent,left,right) describe ONE strictly binary tree whose
// ids are in-order ranks (the property every query relies on).  Iterative.
static int validate_tree(int64_t n, const int32_t *parent, const int32_t *left, const int32_t *right,
                         int32_t *root_out, int64_t *leaves_out) {
    int32_t root = -1;
    int64_t leaves = 0;
    for (int64_t v = 0; v < n; ++v) {
        int32_t p = parent[v], l = left[v], r = right[v];
        if (p < -1 || p >= n || l < -1 || l >= n || r < -1 || r >= n) {
            st_set_error("node %lld: parent/child index out of range", (long long)v);
            return ST_ERR_INVALID_ARG;
        }
        if (p == -1) {
            if (root != -1) {
                st_set_error("more than one root (nodes %d and %lld)", root, (long long)v);
                return ST_ERR_NOT_BINARY;
            }
            root = int32_t(v);
        } else if (left[p] != v && right[p] != v) {
            st_set_error("node %lld: parent %d does not list it as a child", (long long)v, p);
            return ST_ERR_NOT_BINARY;
        }
        if ((l == -1) != (r == -1)) {
            st_set_error("node %lld has exactly one child (tree must be strictly bifurcating)",
                         (long long)v);
            return ST_ERR_NOT_BINARY;
        }
        if (l == -1) {
            ++leaves;
        } else {
            if (l == r || parent[l] != v || parent[r] != v) {
                st_set_error("node %lld: children do not point back to it", (long long)v);
                return ST_ERR_NOT_BINARY;
            }
        }
    }
    if (root == -1) {
        st_set_error("no root (no node with parent -1)");
        return ST_ERR_NOT_BINARY;
    }
    // in-order walk without recursion or a stack: Morris-free, parent pointers suffice
    int64_t next_id = 0;
    int32_t v = root;
    while (left[v] != -1) v = left[v];
    for (;;) {
        if (v != next_id) {
            st_set_error("ids are not in-order ranks: in-order position %lld holds node %d",
                         (long long)next_id, v);
            return ST_ERR_NOT_INORDER;
        }
        ++next_id;
        if (right[v] != -1) {
            v = right[v];
            while (left[v] != -1) v = left[v];
        } else {
            int32_t c = v;
            v = parent[v];
            while (v != -1 && right[v] == c) {
                c = v;
                v = parent[v];
            }
            if (v == -1) break;
        }
        if (next_id > n) break;
    }
    if (next_id != n) {
        st_set_error("tree is not connected: in-order walk visited %lld of %lld nodes",
                     (long long)next_id, (long long)n);
        return ST_ERR_NOT_BINARY;
    }
    *root_out = root;
    *leaves_out = leaves;
    return ST_OK;
}

// ------------------------------------------------------------ create --------
// Upper bound on RMQ blocks: the block tables (8 + 2*levels bytes per block) are
// copied to shared memory by every query CTA.  1024 blocks x 10 levels -> <= 96 KB
// (two 512-thread CTAs per SM still fit).
// SUCHTREE_B200_MAX_BLOCKS overrides (power of two, <= 4096) for experiments.
static int st_max_blocks() {
    int v = 1024;
    if (const char *e = getenv("SUCHTREE_B200_MAX_BLOCKS")) {
        int x = atoi(e);
        if (x >= 1 && x <= 4096) v = x;
    }
    return v;
}
static const int ST_DEFAULT_MICRO_SHIFT = 5;

template <typename T>
static int dev_alloc(T **p, size_t count, int64_t *acc) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), bytes);
    if (e != cudaSuccess) {
        st_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return ST_ERR_NOMEM;
    }
    if (acc) *acc += int64_t(bytes);
    return ST_OK;
}

extern "C" void st_tree_destroy(st_tree *t) {
    if (!t) return;
    DeviceGuard g(t->device);
    cudaFree(t->d_rec);
    cudaFree(t->d_rec16_base);
    cudaFree(t->d_depth);
    if (t->d_tables) {
        cudaFree(t->d_tables);  // d_stk / d_brd / d_stk32 / d_brd8 / d_bid point into it
    } else {                    // creation failed before the blob was packed
        cudaFree(t->d_stk32);
        cudaFree(t->d_brd8);
        cudaFree(t->d_bid);
        cudaFree(t->d_stk);
        cudaFree(t->d_brd);
    }
    cudaFree(t->d_mst);
    cudaFree(t->d_status);
    delete t;
}

extern "C" int st_tree_create(int device, int64_t n_nodes, const int32_t *parent,
                              const int32_t *left, const int32_t *right, const float *edge_len,
                              int block_shift, int micro_shift, st_tree **out) {
    return st_tree_create_ex(device, n_nodes, parent, left, right, edge_len, block_shift, micro_shift, 0,
                             out);
}

extern "C" int st_tree_create_ex(int device, int64_t n_nodes, const int32_t *parent,
                                 const int32_t *left, const int32_t *right, const float *edge_len,
                                 int block_shift, int micro_shift, int flags, st_tree **out) {
    if (!out) return ST_ERR_INVALID_ARG;
    *out = nullptr;
    if (!parent || !left || !right || !edge_len || n_nodes < 1) {
        st_set_error("st_tree_create: NULL array or n_nodes < 1");
        return ST_ERR_INVALID_ARG;
    }
    if (n_nodes >= (int64_t(1) << 31) - 1) {
        st_set_error("st_tree_create: n_nodes must be < 2^31-1");
        return ST_ERR_INVALID_ARG;
    }
    int32_t root = -1;
    int64_t n_leaves = 0;
    int rc = validate_tree(n_nodes, parent, left, right, &root, &n_leaves);
    if (rc != ST_OK) return rc;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        st_set_error("no CUDA device available (this library has no CPU fallback)");
        return ST_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        st_set_error("device %d out of range (%d devices)", device, ndev);
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        st_set_error("cudaSetDevice(%d) failed", device);
        return ST_ERR_CUDA;
    }

    st_tree *t = new (std::nothrow) st_tree();
    if (!t) return ST_ERR_NOMEM;
    t->device = device;
    t->n_nodes = n_nodes;
    t->n_leaves = n_leaves;
    t->root = root;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete t;
        st_set_error("cudaGetDeviceProperties failed");
        return ST_ERR_CUDA;
    }
    t->sm_count = prop.multiProcessorCount;
    {   // stream-ordered scratch (matrix set-up tables) should be recycled, not returned
        // to the driver at every synchronisation
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = uint64_t(1) << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }

    // block geometry: smallest power-of-two block (>= micro block) with <= ST_MAX_BLOCKS blocks
    const int32_t n = int32_t(n_nodes);
    int ms = micro_shift > 0 ? micro_shift : ST_DEFAULT_MICRO_SHIFT;
    if (ms > 10) ms = 10;
    int bs = block_shift > 0 ? block_shift : ms;
    if (bs < ms) bs = ms;
    const int max_blocks = st_max_blocks();
    while (((int64_t(n) + (int64_t(1) << bs) - 1) >> bs) > max_blocks) ++bs;
    t->block_shift = bs;
    t->micro_shift = ms;
    t->n_blocks = int32_t((int64_t(n) + (int64_t(1) << bs) - 1) >> bs);
    t->n_micro = int32_t((int64_t(n) + (int64_t(1) << ms) - 1) >> ms);
    // level k answers spans of [2^k, 2^(k+1)) full blocks; the longest span is n_blocks-2
    t->st_levels = 1;
    while ((1 << t->st_levels) <= std::max(t->n_blocks - 2, 1)) ++t->st_levels;
    t->m_levels = std::max(bs - ms, 1);  // spans of < 2^(bs-ms) micro blocks inside one block

#define ST_TRY(x)              \
    do {                       \
        int _rc = (x);         \
        if (_rc != ST_OK) {    \
            st_tree_destroy(t);\
            return _rc;        \
        }                      \
    } while (0)
#define ST_TRY_CUDA(call)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            st_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,\
                         __LINE__);                                                       \
            st_tree_destroy(t);                                                           \
            return ST_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

    ST_TRY(dev_alloc(&t->d_rec, size_t(n), &t->index_bytes));
    ST_TRY(dev_alloc(&t->d_depth, size_t(n), &t->index_bytes));
    ST_TRY(dev_alloc(&t->d_stk, size_t(t->st_levels) * t->n_blocks, &t->index_bytes));
    ST_TRY(dev_alloc(&t->d_brd, size_t(t->n_blocks), &t->index_bytes));
    ST_TRY(dev_alloc(&t->d_mst, size_t(t->m_levels) * t->n_micro, &t->index_bytes));
    ST_TRY(dev_alloc(&t->d_status, 1, &t->index_bytes));
    ST_TRY_CUDA(cudaMemset(t->d_status, 0, sizeof(RangeStatus)));

    // ---- temporaries for the build
    int32_t *d_parent = nullptr, *d_anc[2] = {nullptr, nullptr}, *d_dep[2] = {nullptr, nullptr};
    float *d_edge = nullptr;
    double *d_rhi[2] = {nullptr, nullptr}, *d_rlo[2] = {nullptr, nullptr};
    int *d_flag = nullptr;
    auto free_tmp = [&]() {
        cudaFree(d_parent);
        cudaFree(d_edge);
        cudaFree(d_flag);
        for (int i = 0; i < 2; ++i) {
            cudaFree(d_anc[i]);
            if (d_dep[i] != t->d_depth) cudaFree(d_dep[i]);
            cudaFree(d_rhi[i]);
            cudaFree(d_rlo[i]);
        }
    };
#define ST_TRY2(x)             \
    do {                       \
        int _rc = (x);         \
        if (_rc != ST_OK) {    \
            free_tmp();        \
            st_tree_destroy(t);\
            return _rc;        \
        }                      \
    } while (0)
#define ST_TRY2_CUDA(call)                                                                \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            st_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,\
                         __LINE__);                                                       \
            free_tmp();                                                                   \
            st_tree_destroy(t);                                                           \
            return ST_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

    ST_TRY2(dev_alloc(&d_parent, size_t(n), nullptr));
    ST_TRY2(dev_alloc(&d_edge, size_t(n), nullptr));
    ST_TRY2(dev_alloc(&d_flag, 2, nullptr));
    for (int i = 0; i < 2; ++i) {
        ST_TRY2(dev_alloc(&d_anc[i], size_t(n), nullptr));
        ST_TRY2(dev_alloc(&d_dep[i], size_t(n), nullptr));
        ST_TRY2(dev_alloc(&d_rhi[i], size_t(n), nullptr));
        ST_TRY2(dev_alloc(&d_rlo[i], size_t(n), nullptr));
    }
    cudaStream_t s = nullptr;  // legacy default stream: build is synchronous by design
    ST_TRY2_CUDA(cudaMemcpy(d_parent, parent, size_t(n) * 4, cudaMemcpyHostToDevice));
    ST_TRY2_CUDA(cudaMemcpy(d_edge, edge_len, size_t(n) * 4, cudaMemcpyHostToDevice));

    const int TPB = 256;
    const int grid = (n + TPB - 1) / TPB;
    k_jump_init<<<grid, TPB, 0, s>>>(n, d_parent, d_edge, d_anc[0], d_dep[0], d_rhi[0], d_rlo[0]);
    int cur = 0;
    for (int round = 0; round < 40; ++round) {
        ST_TRY2_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
        k_jump_round<<<grid, TPB, 0, s>>>(n, d_anc[cur], d_dep[cur], d_rhi[cur], d_rlo[cur],
                                          d_anc[cur ^ 1], d_dep[cur ^ 1], d_rhi[cur ^ 1],
                                          d_rlo[cur ^ 1], d_flag);
        cur ^= 1;
        int live = 0;
        ST_TRY2_CUDA(cudaMemcpy(&live, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (!live) break;
    }
    // after the loop every anc is -1 except possibly one more hop: run until quiescent
    // (the flag is raised only when a second-level ancestor exists, so one extra
    //  round has already folded the last hop in).
    ST_TRY2_CUDA(cudaMemcpyAsync(t->d_depth, d_dep[cur], size_t(n) * 4, cudaMemcpyDeviceToDevice, s));
    ST_TRY2_CUDA(cudaMemsetAsync(d_flag + 1, 0, sizeof(int), s));
    k_max_depth<<<grid, TPB, 0, s>>>(n, t->d_depth, d_flag + 1);
    int maxd = 0;
    ST_TRY2_CUDA(cudaMemcpy(&maxd, d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost));
    t->depth = maxd + 1;  // reference counts nodes on the path, MuchTree.pyx:218-225

    {
        int T = std::min(256, 1 << bs);
        k_block_records<<<t->n_blocks, T, 0, s>>>(n, bs, t->d_depth, d_rhi[cur], d_rlo[cur],
                                                  t->d_rec, t->d_stk, t->d_brd);
        k_block_sparse_table<<<1, 1024, 0, s>>>(t->n_blocks, t->st_levels, t->d_stk);
        int gm = (t->n_micro + TPB - 1) / TPB;
        k_micro_min<<<gm, TPB, 0, s>>>(n, ms, t->n_micro, t->d_depth, t->d_mst);
        for (int k = 1; k < t->m_levels; ++k)
            k_micro_level<<<gm, TPB, 0, s>>>(t->n_micro, 1 << (k - 1),
                                             t->d_mst + size_t(k - 1) * t->n_micro,
                                             t->d_mst + size_t(k) * t->n_micro);
    }
    ST_TRY2_CUDA(cudaGetLastError());

    // ---- compact layout when it loses nothing: every root distance exact in fp64
    //      (all low words zero) and (depth, offset) / (depth, block) fit 32 bits
    int ts = 1;
    while ((1 << ts) < t->n_blocks) ++ts;
    bool want_compact = !(flags & ST_TREE_WIDE_LAYOUT);
    if (const char *e = getenv("SUCHTREE_B200_LAYOUT")) want_compact = want_compact && e[0] != 'w';
    if (want_compact && (uint64_t(maxd) >> (32 - std::max(bs, ts))) == 0) {
        // the depths fit: 32-bit block tables in any case ...
        int64_t cbytes = 0;
        ST_TRY2(dev_alloc(&t->d_stk32, size_t(t->st_levels) * t->n_blocks, &cbytes));
        ST_TRY2(dev_alloc(&t->d_brd8, size_t(t->n_blocks), &cbytes));
        ST_TRY2(dev_alloc(&t->d_bid, size_t(t->n_blocks), &cbytes));
        const int tot = t->st_levels * t->n_blocks;
        k_compact_tables<<<(tot + TPB - 1) / TPB, TPB, 0, s>>>(t->n_blocks, t->st_levels, bs, ts, t->d_stk,
                                                              t->d_brd, t->d_stk32, t->d_brd8, t->d_bid);
        ST_TRY2_CUDA(cudaGetLastError());
        t->compact_tables = 1;
        t->index_bytes += cbytes - int64_t(size_t(t->st_levels) * t->n_blocks * 8);
        // ... and 16-byte records when every root distance is exact in fp64
        ST_TRY2_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
        k_any_nonzero<<<grid, TPB, 0, s>>>(n, d_rlo[cur], d_flag);
        int inexact = 0;
        ST_TRY2_CUDA(cudaMemcpy(&inexact, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (const char *e = getenv("SUCHTREE_B200_LAYOUT"))
            if (e[0] == 't') inexact = 1;  // "tables": keep the wide records (experiments)
        if (!inexact) {
            // Two 16-byte records share a 32-byte sector.  A leaf's parent is its in-order
            // neighbour (id - 1 or id + 1): slot = id + shift with the shift that puts most leaves
            // into the SAME sector as their parent, so that a query whose MRCA is the parent of
            // one endpoint (ladder-like trees: always) finds rd[mrca] in a sector it loads anyway
            // (st_ld_rec_c<true>).  n + 2 slots, zero padded: the paired load never leaves the array.
            int64_t votes[2] = {0, 0};
            for (int64_t v = 0; v < n_nodes; v += 2)
                if (left[v] == -1 && parent[v] >= 0) ++votes[parent[v] == v + 1 ? 0 : 1];
            t->rec16_shift = votes[1] > votes[0] ? 1 : 0;
            if (const char *e = getenv("SUCHTREE_B200_REC_SHIFT")) t->rec16_shift = atoi(e) & 1;
            int64_t rbytes = 0;
            ST_TRY2(dev_alloc(&t->d_rec16_base, size_t(n) + 2, &rbytes));
            ST_TRY2_CUDA(cudaMemsetAsync(t->d_rec16_base, 0, (size_t(n) + 2) * sizeof(NodeRec16), s));
            t->d_rec16 = t->d_rec16_base + t->rec16_shift;
            k_compact_records<<<grid, TPB, 0, s>>>(n, bs, t->d_rec, t->d_rec16);
            ST_TRY2_CUDA(cudaGetLastError());
            ST_TRY2_CUDA(cudaDeviceSynchronize());
            t->index_bytes += rbytes - int64_t(size_t(n) * sizeof(NodeRec)) - int64_t(size_t(t->n_blocks) * 16);
            cudaFree(t->d_rec); t->d_rec = nullptr;
            cudaFree(t->d_brd); t->d_brd = nullptr;
            t->compact = 1;
        } else {
            t->index_bytes -= int64_t(size_t(t->n_blocks) * 8);  // brd8 is not used with wide records
        }
        ST_TRY2_CUDA(cudaDeviceSynchronize());
        cudaFree(t->d_stk); t->d_stk = nullptr;
    }
    ST_TRY2_CUDA(cudaDeviceSynchronize());
    free_tmp();

    // ---- the block tables as ONE blob in their shared-memory layout: every query CTA stages
    //      it with a single TMA bulk copy; the view's table pointers point into it
    const int tmode = t->compact ? 1 : (t->compact_tables ? 2 : 0);
    const TableLayout TL = st_table_layout(t->n_blocks, t->st_levels, tmode);
    {
        int64_t blob_bytes = 0;
        ST_TRY(dev_alloc(&t->d_tables, size_t(TL.bytes), &blob_bytes));  // counted below
        ST_TRY_CUDA(cudaMemset(t->d_tables, 0, size_t(TL.bytes)));
        const size_t nb = size_t(t->n_blocks), lv = size_t(t->st_levels);
        if (tmode == 1) ST_TRY_CUDA(cudaMemcpy(t->d_tables + TL.off_brd, t->d_brd8, nb * 8, cudaMemcpyDeviceToDevice));
        else ST_TRY_CUDA(cudaMemcpy(t->d_tables + TL.off_brd, t->d_brd, nb * 16, cudaMemcpyDeviceToDevice));
        if (tmode == 0) {
            ST_TRY_CUDA(cudaMemcpy(t->d_tables + TL.off_stk, t->d_stk, lv * nb * 8, cudaMemcpyDeviceToDevice));
        } else {
            ST_TRY_CUDA(cudaMemcpy(t->d_tables + TL.off_bid, t->d_bid, nb * 4, cudaMemcpyDeviceToDevice));
            ST_TRY_CUDA(cudaMemcpy(t->d_tables + TL.off_stk, t->d_stk32, lv * nb * 4, cudaMemcpyDeviceToDevice));
        }
        // the separately built arrays are now dead
        t->index_bytes += blob_bytes - int64_t(tmode == 0 ? lv * nb * 8 + nb * 16
                                                          : lv * nb * 4 + nb * 4 + (tmode == 1 ? nb * 8 : nb * 16));
        cudaFree(t->d_stk); cudaFree(t->d_brd); cudaFree(t->d_stk32); cudaFree(t->d_brd8); cudaFree(t->d_bid);
        t->d_stk = nullptr; t->d_brd = nullptr; t->d_stk32 = nullptr; t->d_brd8 = nullptr; t->d_bid = nullptr;
        if (tmode == 1) t->d_brd8 = reinterpret_cast<double *>(t->d_tables + TL.off_brd);
        else t->d_brd = reinterpret_cast<double2 *>(t->d_tables + TL.off_brd);
        if (tmode == 0) t->d_stk = reinterpret_cast<uint64_t *>(t->d_tables + TL.off_stk);
        else {
            t->d_bid = reinterpret_cast<int32_t *>(t->d_tables + TL.off_bid);
            t->d_stk32 = reinterpret_cast<uint32_t *>(t->d_tables + TL.off_stk);
        }
    }
    t->query_smem_bytes = TL.bytes;
    t->view.tables = t->d_tables;
    t->view.tables_bytes = TL.bytes;
    t->view.rec16 = t->d_rec16;
    t->view.stk32 = t->d_stk32;
    t->view.brd8 = t->d_brd8;
    t->view.bid = t->d_bid;
    t->view.compact = t->compact;
    t->view.compact_tables = t->compact_tables;
    t->view.table_shift = ts;
    t->view.rec = t->d_rec;
    t->view.depth = t->d_depth;
    t->view.stk = t->d_stk;
    t->view.brd = t->d_brd;
    t->view.mst = t->d_mst;
    t->view.status = t->d_status;
    t->view.n_nodes = n;
    t->view.id_bits = std::max(1, st_ceil_log2_i64(n));
    t->view.n_blocks = t->n_blocks;
    t->view.n_micro = t->n_micro;
    t->view.block_shift = bs;
    t->view.micro_shift = ms;
    t->view.st_levels = t->st_levels;
    t->view.m_levels = t->m_levels;
    if (t->compact && n_leaves > 1) {
        const int32_t probes = 16384;
        unsigned int *d_counts = nullptr, h_counts[2] = {0, 0};
        ST_TRY_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_counts), 2 * sizeof(unsigned int)));
        cudaMemset(d_counts, 0, 2 * sizeof(unsigned int));
        k_probe_paired<<<probes / 256, 256>>>(t->view, uint32_t(n_leaves), probes, d_counts);
        cudaError_t e = cudaMemcpy(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost);
        cudaFree(d_counts);
        if (e != cudaSuccess) {
            st_set_error("paired-record probe failed: %s", cudaGetErrorString(e));
            st_tree_destroy(t);
            return ST_ERR_CUDA;
        }
        t->probe_third_gather = double(h_counts[0]) / probes;
        t->probe_neighbour_hit = double(h_counts[1]) / probes;
        // the paired kernel costs 2-6 % where it never helps and gains 25 % where it always does
        t->paired = t->probe_neighbour_hit > 0.2 ? 1 : 0;
    }
    st_hostctx_prewarm(device);  // staging of the host-buffer entry points: once per device, in the background
    *out = t;
    return ST_OK;
}

extern "C" int st_tree_get_info(const st_tree *t, st_tree_info *info) {
    if (!t || !info) return ST_ERR_INVALID_ARG;
    info->n_nodes = t->n_nodes;
    info->n_leaves = t->n_leaves;
    info->root = t->root;
    info->depth = t->depth;
    info->device = t->device;
    info->block_shift = t->block_shift;
    info->micro_shift = t->micro_shift;
    info->n_blocks = t->n_blocks;
    info->index_bytes = t->index_bytes;
    info->query_smem_bytes = t->query_smem_bytes;
    info->sm_count = t->sm_count;
    info->layout = t->compact ? 1 : (t->compact_tables ? 2 : 0);
    info->paired_records = t->paired;
    info->probe_third_gather = t->probe_third_gather;
    info->probe_neighbour_hit = t->probe_neighbour_hit;
    return ST_OK;
}

__global__ void k_export_rd(const TreeView tv, double *__restrict__ hi, double *__restrict__ lo) {
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= tv.n_nodes) return;
    const dd r = st_ld_rd(tv, v);
    if (hi) hi[v] = r.hi;
    if (lo) lo[v] = r.lo;
}

extern "C" int st_tree_export(const st_tree *t, int32_t *depth, double *rd_hi, double *rd_lo) {
    if (!t) return ST_ERR_INVALID_ARG;
    DeviceGuard g(t->device);
    const int32_t n = int32_t(t->n_nodes);
    if (depth) ST_CUDA(cudaMemcpy(depth, t->d_depth, size_t(n) * 4, cudaMemcpyDeviceToHost));
    if (rd_hi || rd_lo) {
        double *d_hi = nullptr, *d_lo = nullptr;
        ST_CUDA(cudaMalloc(&d_hi, size_t(n) * 8));
        if (cudaMalloc(&d_lo, size_t(n) * 8) != cudaSuccess) {
            cudaFree(d_hi);
            st_set_error("cudaMalloc failed in st_tree_export");
            return ST_ERR_NOMEM;
        }
        k_export_rd<<<(n + 255) / 256, 256>>>(t->view, d_hi, d_lo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess && rd_hi) e = cudaMemcpy(rd_hi, d_hi, size_t(n) * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && rd_lo) e = cudaMemcpy(rd_lo, d_lo, size_t(n) * 8, cudaMemcpyDeviceToHost);
        cudaFree(d_hi);
        cudaFree(d_lo);
        if (e != cudaSuccess) {
            st_set_error("st_tree_export: %s", cudaGetErrorString(e));
            return ST_ERR_CUDA;
        }
    }
    return ST_OK;
}
