// Device-side building blocks: double-double arithmetic, packed (depth,id) keys,
// the range-minimum MRCA lookup, Philox4x32-10, cache-hinted loads/stores.
#pragma once

#include "st_internal.cuh"

// ---------------------------------------------------------------- keys ------
// MRCA(a,b) = the node of minimum depth among ids [min(a,b), max(a,b)]
// (ids are in-order ranks, MuchTree.pyx:171-180; the minimum is unique).
// A key packs (depth << 32 | id) so that an unsigned 64-bit min is an argmin.
__device__ __forceinline__ uint64_t st_key(int32_t depth, int32_t id) {
    return (uint64_t(uint32_t(depth)) << 32) | uint32_t(id);
}
__device__ __forceinline__ int32_t st_key_id(uint64_t k) { return int32_t(uint32_t(k)); }
__device__ __forceinline__ uint64_t st_min64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// ------------------------------------------------------- double-double ------
// Root distances are kept as unevaluated sums hi + lo so that
// rd[a] + rd[b] - 2 rd[mrca] does not lose the short edges (the reference's
// 2.2e-16 polytomy epsilon, MuchTree.pyx:136) against O(1) root distances.
// Explicit _rn intrinsics: no FMA contraction, no reassociation.
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd dd_two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return dd{s, e};
}
__device__ __forceinline__ dd dd_fast_two_sum(double a, double b) {  // |a| >= |b|
    double s = __dadd_rn(a, b);
    double e = __dsub_rn(b, __dsub_rn(s, a));
    return dd{s, e};
}
__device__ __forceinline__ dd dd_add(dd x, dd y) {  // accurate (IEEE-style) variant
    dd s = dd_two_sum(x.hi, y.hi);
    dd t = dd_two_sum(x.lo, y.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = dd_fast_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return dd_fast_two_sum(s.hi, s.lo);
}
// d = rd[a] + rd[b] - 2 rd[m], rounded once to fp64
__device__ __forceinline__ double st_patristic(dd ra, dd rb, dd rm) {
    dd s = dd_add(ra, rb);
    dd m2{-2.0 * rm.hi, -2.0 * rm.lo};  // exact scaling
    dd r = dd_add(s, m2);
    return r.hi;  // normalised: hi = fl(hi + lo)
}
// The same value for operands whose low words are zero (compact layout: every root distance is
// exact in fp64): st_patristic() with the additions of +0.0 and the renormalisations of already
// normalised pairs taken out -- 18 fp64 operations instead of 42, the same bits (x + 0.0 = x for
// every x but -0.0, which can only arise where the result is 0).  The linked-tree moment kernels
// are issue-bound on exactly these operations (profiles/r02_summary.md).
__device__ __forceinline__ double st_patristic_c(double ra, double rb, double rm) {
    const dd s = dd_two_sum(ra, rb);           // dd_add(ra, rb): exact, already normalised
    const dd t = dd_two_sum(s.hi, -2.0 * rm);  // dd_add(s, -2 rm): high words ...
    const double lo = __dadd_rn(t.lo, s.lo);   // ... low words (the second operand's is zero)
    const dd u = dd_fast_two_sum(t.hi, lo);
    return __dadd_rn(u.hi, u.lo);              // the last renormalisation's high word
}
// one-add variant for pre-combined operands (matrix writer): fl(x + y)
__device__ __forceinline__ double dd_add_to_double(dd x, dd y) {
    dd s = dd_two_sum(x.hi, y.hi);
    double lo = __dadd_rn(__dadd_rn(x.lo, y.lo), s.lo);
    return __dadd_rn(s.hi, lo);
}

// ------------------------------------------------------ cache-hinted I/O ----
// pair streams and results are touched once: keep them out of L1 and mark them
// evict-first in L2 so the index (rec[] etc.) stays resident.
__device__ __forceinline__ uint64_t st_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int4 st_ld_stream_int4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(st_policy_evict_first()));
    return r;
}
// 256-bit streaming load / stores (sm_100+)
__device__ __forceinline__ void st_ld_stream_256(const void *p, uint64_t (&w)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
                 : "l"(p), "l"(st_policy_evict_first()));
}
__device__ __forceinline__ void st_st_stream_f64x4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "d"(a), "d"(b), "d"(c), "d"(d), "l"(st_policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_st_stream_i32x4(int32_t *p, int32_t a, int32_t b, int32_t c, int32_t d) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "r"(a), "r"(b), "r"(c), "r"(d), "l"(st_policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_st_stream_f64x2(double *p, double a, double b) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p),
                 "d"(a), "d"(b), "l"(st_policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_st_stream_f64(double *p, double a) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(a),
                 "l"(st_policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_st_stream_i32x2(int32_t *p, int32_t a, int32_t b) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.s32 [%0], {%1,%2}, %3;" ::"l"(p),
                 "r"(a), "r"(b), "l"(st_policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_st_stream_i32(int32_t *p, int32_t a) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(a),
                 "l"(st_policy_evict_first())
                 : "memory");
}
// Layout mode of the accessors below:
//   0 = wide records (32 B, double-double rd, 64-bit keys) + wide block tables
//   1 = compact records (16 B) + compact block tables
//   3 = wide records + compact block tables (inexact root distances, ordinary depths)
//   2 = decided at run time from tv.compact / tv.compact_tables (kernels off the headline path)
template <int M>
__device__ __forceinline__ bool st_compact(const TreeView &tv) {  // records
    return M == 2 ? tv.compact != 0 : M == 1;
}
template <int M>
__device__ __forceinline__ bool st_ctab(const TreeView &tv) {  // block tables
    return M == 2 ? tv.compact_tables != 0 : (M == 1 || M == 3);
}

// one node record = one sector (or half of one), fetched with ONE 256-bit load
// (ld.global.v4.b64, sm_100+) or one 128-bit load, marked evict-last so the index
// outlives the streams in L2.  Compact keys are widened to the common form.
struct RecRaw {
    double rd_hi, rd_lo;
    uint64_t suf, pre;
};
template <int M = 2>
__device__ __forceinline__ RecRaw st_ld_rec(const TreeView &tv, int32_t id) {
    RecRaw r;
    uint64_t a, b;
    if (st_compact<M>(tv)) {
        asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(tv.rec16 + id));
        const uint32_t mask = (1u << tv.block_shift) - 1u, base = uint32_t(id) & ~mask;
        const uint32_t sf = uint32_t(b), pr = uint32_t(b >> 32);
        r.rd_lo = 0.0;
        r.suf = (uint64_t(sf >> tv.block_shift) << 32) | (base + (sf & mask));
        r.pre = (uint64_t(pr >> tv.block_shift) << 32) | (base + (pr & mask));
    } else {
        asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(a), "=l"(b), "=l"(r.suf), "=l"(r.pre)
                     : "l"(tv.rec + id));
        r.rd_lo = __longlong_as_double((long long)b);
    }
    r.rd_hi = __longlong_as_double((long long)a);
    return r;
}
template <int M = 2>
__device__ __forceinline__ dd st_ld_rd(const TreeView &tv, int32_t id) {
    if (st_compact<M>(tv)) return dd{__ldg(&tv.rec16[id].rd), 0.0};
    double2 a = __ldg(reinterpret_cast<const double2 *>(tv.rec + id));
    return dd{a.x, a.y};
}

// ------------------------------------------------------------- RMQ ----------
// Block-level tables (shared memory in the query kernels, global memory in the
// matrix kernels' few set-up queries): sparse table of keys over blocks and the
// root distance of every block minimum.
struct SmemTables {
    const uint64_t *stk;   // wide: [levels][n_blocks] (depth << 32 | id)
    const double2 *brd;    // wide: [n_blocks]
    const uint32_t *stk32; // compact: [levels][n_blocks] (depth << table_shift | block)
    const double *brd8;    // compact: [n_blocks]
    const int32_t *bid;    // compact: [n_blocks]
};
__device__ __forceinline__ SmemTables st_global_tables(const TreeView &tv) {
    return SmemTables{tv.stk, tv.brd, tv.stk32, tv.brd8, tv.bid};  // pointers into the device blob
}

__device__ __forceinline__ uint64_t st_scan_depth(const int32_t *__restrict__ depth, int32_t s,
                                                  int32_t e) {
    uint64_t best = ~0ull;
    for (int32_t i = s; i <= e; ++i) best = st_min64(best, st_key(__ldg(depth + i), i));
    return best;
}

// slow path: lo and hi in the same block -> micro blocks + per-level micro table.
// Kept out of line (it is rare for far-apart pairs) and fed scalars only, so
// the fast path holds no stack frame.
static __device__ __noinline__ uint64_t st_rmq_inblock(const int32_t *__restrict__ depth,
                                                       const uint64_t *__restrict__ mst,
                                                       int32_t n_micro, int ms, int32_t lo,
                                                       int32_t hi) {
    int32_t mlo = lo >> ms, mhi = hi >> ms;
    if (mlo == mhi) return st_scan_depth(depth, lo, hi);
    uint64_t best = st_min64(st_scan_depth(depth, lo, ((mlo + 1) << ms) - 1),
                             st_scan_depth(depth, mhi << ms, hi));
    int32_t span = mhi - mlo - 1;
    if (span > 0) {
        int k = 31 - __clz(span);
        const uint64_t *lvl = mst + size_t(k) * n_micro;
        best = st_min64(best, st_min64(__ldg(lvl + mlo + 1), __ldg(lvl + mhi - (1 << k))));
    }
    return best;
}

// key of the MRCA given both endpoint records (already loaded).  *from_table is set
// when the winner is a block minimum taken from the block table: its root distance
// and id then come from the table (st_mrca_rd / st_mrca_id) and no third gather is
// needed.  The low word of the returned key is the node id, except for a compact
// table winner, where it is the BLOCK index.
template <int M = 2>
__device__ __forceinline__ uint64_t st_rmq(const TreeView &tv, const SmemTables &sm, int32_t lo,
                                           int32_t hi, uint64_t suf_lo, uint64_t pre_hi,
                                           bool *from_table) {
    int32_t blo = lo >> tv.block_shift, bhi = hi >> tv.block_shift;
    *from_table = false;
    if (blo == bhi) return st_rmq_inblock(tv.depth, tv.mst, tv.n_micro, tv.micro_shift, lo, hi);
    uint64_t best = st_min64(suf_lo, pre_hi);
    int32_t span = bhi - blo - 1;
    if (span > 0) {
        int k = 31 - __clz(span);
        if (st_ctab<M>(tv)) {
            const uint32_t *lvl = sm.stk32 + k * tv.n_blocks;
            const uint32_t mid = min(lvl[blo + 1], lvl[bhi - (1 << k)]);
            // candidates are distinct nodes and the minimum depth is unique: no ties
            if ((mid >> tv.table_shift) < uint32_t(best >> 32)) {
                best = (uint64_t(mid >> tv.table_shift) << 32) | (mid & ((1u << tv.table_shift) - 1u));
                *from_table = true;
            }
        } else {
            const uint64_t *lvl = sm.stk + k * tv.n_blocks;
            uint64_t mid = st_min64(lvl[blo + 1], lvl[bhi - (1 << k)]);
            if (mid < best) {
                best = mid;
                *from_table = true;
            }
        }
    }
    return best;
}

// root distance of the MRCA: table copy for block minima, one gather otherwise
template <int M = 2>
__device__ __forceinline__ dd st_mrca_rd(const TreeView &tv, const SmemTables &sm, uint64_t key,
                                         bool from_table) {
    int32_t id = st_key_id(key);
    if (from_table) {
        if (st_compact<M>(tv)) return dd{sm.brd8[id], 0.0};
        // compact tables name the BLOCK, wide ones the node
        double2 r = sm.brd[st_ctab<M>(tv) ? id : id >> tv.block_shift];
        return dd{r.x, r.y};
    }
    return st_ld_rd<M>(tv, id);
}
// node id of the MRCA
template <int M = 2>
__device__ __forceinline__ int32_t st_mrca_id(const TreeView &tv, const SmemTables &sm, uint64_t key,
                                              bool from_table) {
    if (from_table && st_ctab<M>(tv)) return sm.bid[st_key_id(key)];
    return st_key_id(key);
}

// one (a, b) query, ids ordered
struct PairQ {
    int32_t lo, hi;
    bool bad;
};


// Lean form for the compact layout (PR = 2): per endpoint only what the pair needs -- its root
// distance and ONE 32-bit key (suffix key of lo, prefix key of hi) -- 3 registers per record
// instead of 8, 32-bit depth compares instead of widened 64-bit keys.  Same results bit for bit.
struct RecC {
    double rd;
    uint32_t key;
    double nb_rd;  // NB only: root distance of the node in the other half of the 32-byte sector
};
// NB = true (paired records): ONE 256-bit load of the 32-byte sector that holds the node's record
// AND its slot neighbour's (slot = id + shift, see st_tree_create): the neighbour of most leaves is
// their parent, so when the MRCA turns out to be that node its root distance is already here -- no
// third, dependent gather (ladder-like trees, sister queries).  NB = false: the 16-byte record alone
template <bool NB>
__device__ __forceinline__ RecC st_ld_rec_c(const TreeView &tv, int32_t id, bool hi_side) {
    if (NB) {
        const char *p = reinterpret_cast<const char *>(tv.rec16 + id);
        const bool upper = (reinterpret_cast<uintptr_t>(p) & 16) != 0;
        uint64_t w0, w1, w2, w3;
        asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3)
                     : "l"(p - (upper ? 16 : 0)));
        const uint64_t keys = upper ? w3 : w1;
        return RecC{__longlong_as_double((long long)(upper ? w2 : w0)), hi_side ? uint32_t(keys >> 32) : uint32_t(keys),
                    __longlong_as_double((long long)(upper ? w0 : w2))};
    }
    uint64_t a, b;
    asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(tv.rec16 + id));
    return RecC{__longlong_as_double((long long)a), hi_side ? uint32_t(b >> 32) : uint32_t(b), 0.0};
}
template <bool NB>
__device__ __forceinline__ void st_pair_c(const TreeView &tv, const SmemTables &sm, const PairQ &q,
                                          const RecC &l, const RecC &h, bool want_d, bool want_m, double &d,
                                          int32_t &m) {
    const int bs = tv.block_shift;
    double rm;
    int32_t id;
    if (q.lo == q.hi) {  // MRCA(a,a) = a
        rm = l.rd;
        id = q.lo;
    } else {
        const int32_t blo = q.lo >> bs, bhi = q.hi >> bs;
        bool ft = false;
        uint32_t mid = 0;
        if (blo == bhi) {
            id = st_key_id(st_rmq_inblock(tv.depth, tv.mst, tv.n_micro, tv.micro_shift, q.lo, q.hi));
        } else {
            const uint32_t ds = l.key >> bs, dp = h.key >> bs;
            const int32_t span = bhi - blo - 1;
            if (span > 0) {
                const int k = 31 - __clz(span);
                const uint32_t *lvl = sm.stk32 + k * tv.n_blocks;
                mid = min(lvl[blo + 1], lvl[bhi - (1 << k)]);
                // candidates are distinct nodes and the minimum depth is unique: no ties
                ft = (mid >> tv.table_shift) < min(ds, dp);
            }
            const uint32_t mask = (1u << bs) - 1u;
            id = ds < dp ? int32_t((uint32_t(q.lo) & ~mask) + (l.key & mask))
                         : int32_t((uint32_t(q.hi) & ~mask) + (h.key & mask));
        }
        if (ft) {  // a block minimum: root distance (and id) from the shared-memory tables
            const uint32_t blk = mid & ((1u << tv.table_shift) - 1u);
            rm = sm.brd8[blk];
            id = want_m ? sm.bid[blk] : 0;
        } else if (NB) {
            // slot parity of an id: which half of its sector, hence which neighbour came along
            const int32_t lo_nb = q.lo + ((reinterpret_cast<uintptr_t>(tv.rec16 + q.lo) & 16) ? -1 : 1);
            const int32_t hi_nb = q.hi + ((reinterpret_cast<uintptr_t>(tv.rec16 + q.hi) & 16) ? -1 : 1);
            if (id == hi_nb) rm = h.nb_rd;
            else if (id == lo_nb) rm = l.nb_rd;
            else if (id == q.lo) rm = l.rd;
            else if (id == q.hi) rm = h.rd;
            else rm = __ldg(&tv.rec16[id].rd);
        } else {
            rm = id == q.lo ? l.rd : (id == q.hi ? h.rd : __ldg(&tv.rec16[id].rd));
        }
    }
    if (want_d) d = st_patristic_c(l.rd, h.rd, rm);
    if (want_m) m = id;
}

// Layout of the block-table blob (device memory AND shared memory: the blob is staged
// into shared memory by one bulk copy).  mode: 0 = wide tables, 1 = compact tables +
// compact records, 2 = compact tables + wide records.  Sections are 16-byte aligned.
struct TableLayout {
    int off_brd, off_bid, off_stk, bytes;
};
__host__ __device__ __forceinline__ TableLayout st_table_layout(int nb, int levels, int mode) {
    TableLayout L;
    L.off_brd = 0;
    const int brd_bytes = (mode == 1 ? nb * 8 : nb * 16);
    L.off_bid = (brd_bytes + 15) & ~15;
    L.off_stk = mode == 0 ? L.off_bid : L.off_bid + ((nb * 4 + 15) & ~15);
    L.bytes = L.off_stk + ((levels * nb * (mode == 0 ? 8 : 4) + 15) & ~15);
    return L;
}
__host__ __device__ __forceinline__ int st_table_bytes(int n_blocks, int st_levels, int mode) {
    return st_table_layout(n_blocks, st_levels, mode).bytes;
}
__host__ __device__ __forceinline__ int st_table_mode(const TreeView &tv) {
    return tv.compact ? 1 : (tv.compact_tables ? 2 : 0);
}
__device__ __forceinline__ SmemTables st_tables_at(const unsigned char *base, const TreeView &tv, int mode) {
    const TableLayout L = st_table_layout(tv.n_blocks, tv.st_levels, mode);
    SmemTables sm{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (mode == 1) sm.brd8 = reinterpret_cast<const double *>(base + L.off_brd);
    else sm.brd = reinterpret_cast<const double2 *>(base + L.off_brd);
    if (mode == 0) {
        sm.stk = reinterpret_cast<const uint64_t *>(base + L.off_stk);
    } else {
        sm.bid = reinterpret_cast<const int32_t *>(base + L.off_bid);
        sm.stk32 = reinterpret_cast<const uint32_t *>(base + L.off_stk);
    }
    return sm;
}

// Stage the block tables into dynamic shared memory (16-byte aligned) with ONE TMA bulk
// copy (cp.async.bulk global -> shared, completion on an mbarrier) issued by thread 0.
// `bar` is an 8-byte shared-memory word owned by the caller (one per table set).
__device__ __forceinline__ void st_tables_issue(const TreeView &tv, unsigned char *smem, uint64_t *bar) {
    if (threadIdx.x == 0) {
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(tv.tables_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(tv.tables), "r"(tv.tables_bytes), "r"(bar_a)
                     : "memory");
    }
}
// call after a __syncthreads() that follows st_tables_issue (the barrier must be
// initialised before anyone polls it)
template <int M = 2>
__device__ __forceinline__ SmemTables st_tables_wait(const TreeView &tv, unsigned char *smem, uint64_t *bar) {
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(bar_a), "r"(0)
            : "memory");
    }
    const int mode = st_compact<M>(tv) ? 1 : (st_ctab<M>(tv) ? 2 : 0);
    return st_tables_at(smem, tv, mode);
}
template <int M = 2>
__device__ __forceinline__ SmemTables st_load_tables(const TreeView &tv, unsigned char *smem, uint64_t *bar) {
    st_tables_issue(tv, smem, bar);
    __syncthreads();
    return st_tables_wait<M>(tv, smem, bar);
}

// ------------------------------------------------------------ Philox --------
// Philox4x32-10 (Salmon et al. 2011), counter-based: the i-th draw depends on
// (key, i) only, so any sharding of the sample stream gives the same numbers.
struct Philox4 {
    uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 st_philox4x32_10(uint64_t counter, uint64_t key) {
    uint32_t c0 = uint32_t(counter), c1 = uint32_t(counter >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = uint32_t(key), k1 = uint32_t(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
// unbiased-enough map of a 32-bit draw to [0, n): floor(u * n / 2^32)
__device__ __forceinline__ uint32_t st_bounded(uint32_t u, uint32_t n) { return __umulhi(u, n); }
