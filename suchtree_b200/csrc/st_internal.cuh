// Internal definitions shared by the translation units of libsuchtree_b200.so.
// Nothing here is part of the C ABI (include/suchtree_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "suchtree_b200.h"  // include/ of the repository, or suchtree_b200/include/ when installed

// ---------------------------------------------------------------- errors ----
void st_set_error(const char *fmt, ...);
void st_set_bad_node(int64_t id);

#define ST_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            st_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                        \
            return ST_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// RAII device switch (the library never leaves the caller's current device changed)
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------ device data ---
// One 32-byte record per node = exactly one L2 sector.  A query touches rec[a],
// rec[b] and rec[mrca] -> three sectors.
struct __align__(32) NodeRec {
    double rd_hi, rd_lo;  // root distance as a double-double
    uint64_t suf;         // (depth<<32 | id) of argmin depth over [v, end of v's block]
    uint64_t pre;         // same over [start of v's block, v]
};
static_assert(sizeof(NodeRec) == 32, "NodeRec must be one sector");

// Compact layout (16 B/node), used when every root distance is exact in fp64 (all
// rd_lo == 0) and the depths leave room: suf/pre = (depth << block_shift) | (id of
// the argmin relative to the start of the node's block).  Halves the L2 footprint
// of the index (32 MB instead of 64 MB at 10^6 leaves).
struct __align__(16) NodeRec16 {
    double rd;
    uint32_t suf, pre;
};
static_assert(sizeof(NodeRec16) == 16, "NodeRec16 must be 16 bytes");

// range status written by query kernels (pinned-host mapped would also do; it
// lives in device memory and is read back by st_check_range / host entry points)
struct RangeStatus {
    unsigned long long max_bad;  // max id >= n_nodes seen (0 = none: 0 is always a valid id)
    long long min_bad;           // min id < 0 seen (0 = none)
};

// Device-side view handed to kernels by value.
struct TreeView {
    const NodeRec *rec;        // [n_nodes]   wide layout (NULL when compact)
    const NodeRec16 *rec16;    // [n_nodes]   compact layout (NULL when wide)
    const int32_t *depth;      // [n_nodes] node depth, root = 0
    const uint64_t *stk;       // [st_levels][n_blocks] block sparse table of packed (depth,id) keys;
                               // level k, entry i = min key over blocks [i, i+2^k); level 0 = block minima
    const double2 *brd;        // [n_blocks] root distance (hi, lo) of each block's minimum node
    // compact block tables: key = (depth << table_shift) | block index of the argmin
    const uint32_t *stk32;     // [st_levels][n_blocks]
    const double *brd8;        // [n_blocks] root distance of each block's minimum node
    const int32_t *bid;        // [n_blocks] id of each block's minimum node
    const unsigned char *tables;  // the block tables above as ONE blob in their shared-memory layout
    int32_t tables_bytes;         // (st_table_layout): staged by one TMA bulk copy per query CTA
    const uint64_t *mst;       // [m_levels][n_micro] micro sparse table (packed keys)
    RangeStatus *status;
    int32_t n_nodes;
    int32_t n_blocks;
    int32_t n_micro;
    int32_t block_shift;
    int32_t micro_shift;
    int32_t st_levels;  // levels stored, level k covers 2^k blocks, level 0 = identity (stored)
    int32_t m_levels;
    int32_t compact;         // 1: rec16 (16-byte records) is the live node array
    int32_t compact_tables;  // 1: stk32 / bid (+ brd8, or brd with wide records) are the live block tables
    int32_t table_shift;     // compact block-table keys: depth << table_shift | block
    int32_t id_bits;         // ceil(log2(n_nodes)), >= 1: width of one id in the bit-packed pair stream (st_query.cu)
};

struct st_tree {
    int device = 0;
    int sm_count = 0;
    int64_t n_nodes = 0, n_leaves = 0;
    int32_t root = -1, depth = 0;
    int32_t block_shift = 0, micro_shift = 0, n_blocks = 0, n_micro = 0, st_levels = 0, m_levels = 0;
    int64_t index_bytes = 0;
    // device allocations
    NodeRec *d_rec = nullptr;
    int32_t *d_depth = nullptr;
    uint64_t *d_stk = nullptr;
    double2 *d_brd = nullptr;
    NodeRec16 *d_rec16 = nullptr;       // = d_rec16_base + rec16_shift: record of node v at slot v + shift
    NodeRec16 *d_rec16_base = nullptr;  // allocation: n_nodes + 2 slots, zero padded, 32-byte aligned
    int rec16_shift = 0;                // 0 | 1: chosen so that most leaves share a 32-byte sector with their parent
    int paired = 0;                     // 1: the pair kernel loads whole sectors (record + neighbour), see k_probe_paired
    double probe_third_gather = 0.0, probe_neighbour_hit = 0.0;
    uint32_t *d_stk32 = nullptr;
    double *d_brd8 = nullptr;
    int32_t *d_bid = nullptr;
    unsigned char *d_tables = nullptr;  // blob the view's table pointers point into
    int compact = 0, compact_tables = 0;
    uint64_t *d_mst = nullptr;
    RangeStatus *d_status = nullptr;
    TreeView view{};
    int query_smem_bytes = 0;

    // (host-buffer entry points take their streams, staging and status word from a lane of
    //  the device's host context, st_hostctx.cuh: a tree handle owns no per-call state)
};

// ------------------------------------------------------- internal launches --
// query kernels (st_query.cu)
// status = NULL: the tree's own word (device API, read by st_check_range); host calls pass
// their lane's word
static const int ST_IDX_PACKED = -2;  // idx_bits value: the bit-packed pair stream (internal: host pipeline only)
int st_launch_pairs(const st_tree *t, const void *d_pairs, int idx_bits, int64_t n, double *d_out,
                    int32_t *d_mrca, cudaStream_t stream, RangeStatus *status = nullptr);
int st_read_range_status(const st_tree *t, cudaStream_t stream, bool *bad);

static inline int st_ceil_log2_i64(int64_t x) {
    int k = 0;
    while ((int64_t(1) << k) < x) ++k;
    return k;
}
