// All-pairs patristic distance matrix.
//
// Replaces SuchTree.pairwise_distances (MuchTree.pyx:1082-1124), which builds
// n(n-1)/2 Python tuples, calls distances_bulk and mirrors the result in a Python
// loop.  Here one kernel writes row blocks of the symmetric fp64 matrix.
//
// Tile = TR rows x TC columns of the matrix, one CTA.  When the id list is
// strictly ascending (the default: all leaves, ids 0,2,4,...) and the tile does
// not straddle the diagonal, the MRCA of (lo, hi) with lo on the "low side" of
// the tile and hi on the "high side" is the best of
//     S(lo) = argmin depth[lo .. lowmax]      one RMQ per low-side row/column
//     M     = argmin depth[lowmax .. highmin] one RMQ per tile
//     P(hi) = argmin depth[highmin .. hi]     one RMQ per high-side row/column
// so a tile of TR*TC elements costs TR+TC+1 index queries; every element is one
// 64-bit compare plus ONE double-double add of pre-combined operands:
//     lo side: key = min(S, M), comb = rd[lo] - 2 rd[argmin], plain = rd[lo]
//     hi side: key = P,         comb = rd[hi] - 2 rd[P],      plain = rd[hi]
//     d = lokey <= hikey ? locomb + hiplain : loplain + hicomb
// No per-element gathers: the kernel is bound by the 8 B/element it writes.
// Tiles on the diagonal, and unsorted id lists, take the per-element query path.
#include <algorithm>

#include "st_device.cuh"
#include "st_hostctx.cuh"

static const int MT = 256;       // threads per CTA
static const int TC = 2 * MT;    // columns per tile (two per thread -> 16-byte stores)
static const int TR = 64;        // rows per tile

__device__ __forceinline__ int32_t mat_id(const int32_t *__restrict__ ids, int64_t k) {
    return ids ? __ldg(ids + k) : int32_t(2 * k);  // default: leaf k has id 2k
}

// full query through global-memory tables (set-up pass and diagonal tiles only)
__device__ __forceinline__ uint64_t mat_query(const TreeView &tv, int32_t a, int32_t b) {
    int32_t lo = min(a, b), hi = max(a, b);
    if (lo == hi) return st_key(__ldg(tv.depth + lo), lo);
    RecRaw rl = st_ld_rec(tv, lo), rh = st_ld_rec(tv, hi);
    const SmemTables g = st_global_tables(tv);
    bool ft;
    const uint64_t k = st_rmq(tv, g, lo, hi, rl.suf, rh.pre, &ft);
    return (k & 0xffffffff00000000ull) | uint32_t(st_mrca_id(tv, g, k, ft));  // always (depth, node id)
}

__device__ __forceinline__ dd dd_minus_2x(dd a, dd m) {  // a - 2m
    return dd_add(a, dd{-2.0 * m.hi, -2.0 * m.lo});
}

__device__ __forceinline__ void mat_store2(double *row, int64_t col, int64_t n, double v0, double v1) {
    double *p = row + col;
    if (col + 1 < n) {
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            st_st_stream_f64x2(p, v0, v1);
        } else {
            st_st_stream_f64(p, v0);
            st_st_stream_f64(p + 1, v1);
        }
    } else if (col < n) {
        st_st_stream_f64(p, v0);
    }
}

// ------------------------------------------------------------ set-up pass ---
// Per-call side tables for an ASCENDING id list (built by k_matrix_sides /
// k_matrix_mid, a few queries per row/column, then read coalesced by every tile):
//   rd[k]                      root distance of ids[k]
//   p*/s* [k] = (depth, comb)  depth of the range argmin and comb = rd[k] - 2 rd[argmin] for
//       pcol: argmin depth[ids[c0] .. ids[k]],   c0 = first column of k's column tile
//       scol: argmin depth[ids[k] .. ids[c1-1]], c1-1 = last column of k's column tile
//       prow / srow: the same relative to k's ROW tile (rows are counted from row_begin)
//   mid[rt][ct] = (depth, rd) of argmin depth[lowmax .. highmin] between a row tile and a
//       column tile that do not overlap
struct __align__(32) SideRec {  // 32 bytes
    double comb_hi, comb_lo;
    uint32_t depth;
    uint32_t pad0;
    uint64_t pad1;
};
struct __align__(32) MidRec {  // 32 bytes
    double rd_hi, rd_lo;
    uint32_t depth;
    uint32_t valid;
    uint64_t pad1;
};
struct MatTables {
    const double2 *rd;     // [n]
    const SideRec *pcol;   // [n]
    const SideRec *scol;   // [n]
    const SideRec *prow;   // [rows]  (index k - row_begin)
    const SideRec *srow;   // [rows]
    const MidRec *mid;     // [row tiles][col tiles]
    int32_t n_ct;
};

__device__ __forceinline__ SideRec mat_side(const TreeView &tv, dd rd, int32_t a, int32_t b) {
    const uint64_t k = mat_query(tv, a, b);
    const dd c = dd_minus_2x(rd, st_ld_rd(tv, st_key_id(k)));
    return SideRec{c.hi, c.lo, uint32_t(k >> 32), 0u, 0ull};
}

__global__ void k_matrix_sides(const TreeView tv, const int32_t *__restrict__ ids, int64_t n,
                               int64_t row_begin, int64_t row_end, double2 *__restrict__ rd_out,
                               SideRec *__restrict__ pcol, SideRec *__restrict__ scol,
                               SideRec *__restrict__ prow, SideRec *__restrict__ srow) {
    const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t id = mat_id(ids, k);
    const dd rd = st_ld_rd(tv, id);
    rd_out[k] = make_double2(rd.hi, rd.lo);
    const int64_t c0 = k / TC * TC, c1 = min(c0 + TC, n);
    pcol[k] = mat_side(tv, rd, mat_id(ids, c0), id);
    scol[k] = mat_side(tv, rd, id, mat_id(ids, c1 - 1));
    if (k >= row_begin && k < row_end) {
        const int64_t r0 = row_begin + (k - row_begin) / TR * TR, r1 = min(r0 + TR, row_end);
        prow[k - row_begin] = mat_side(tv, rd, mat_id(ids, r0), id);
        srow[k - row_begin] = mat_side(tv, rd, id, mat_id(ids, r1 - 1));
    }
}

__global__ void k_matrix_mid(const TreeView tv, const int32_t *__restrict__ ids, int64_t n,
                             int64_t row_begin, int64_t row_end, int32_t n_rt, int32_t n_ct,
                             MidRec *__restrict__ mid) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= int64_t(n_rt) * n_ct) return;
    const int64_t rt = q / n_ct, ct = q % n_ct;
    const int64_t r0 = row_begin + rt * TR, r1 = min(r0 + TR, row_end);
    const int64_t c0 = ct * TC, c1 = min(c0 + TC, n);
    MidRec m{0.0, 0.0, 0u, 0u, 0ull};
    int32_t a = -1, b = -1;
    if (r1 <= c0) { a = mat_id(ids, r1 - 1); b = mat_id(ids, c0); }
    else if (c1 <= r0) { a = mat_id(ids, c1 - 1); b = mat_id(ids, r0); }
    if (a >= 0) {
        const uint64_t k = mat_query(tv, a, b);
        const dd r = st_ld_rd(tv, st_key_id(k));
        m = MidRec{r.hi, r.lo, uint32_t(k >> 32), 1u, 0ull};
    }
    mid[q] = m;
}

__device__ __forceinline__ SideRec ld_side(const SideRec *p) {
    uint64_t x, y, z, w;  // one 256-bit load per record
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(p));
    SideRec s;
    s.comb_hi = __longlong_as_double((long long)x);
    s.comb_lo = __longlong_as_double((long long)y);
    s.depth = uint32_t(z);
    s.pad0 = 0;
    s.pad1 = w;  // padding word of the record (keeps the fourth lane of the load "used")
    return s;
}

// low side meets the tile's middle range: the candidate is the shallower of S and M
__device__ __forceinline__ void mat_apply_mid(SideRec &s, dd plain, uint32_t mdepth, dd mrd) {
    if (mdepth < s.depth) {
        const dd c = dd_minus_2x(plain, mrd);
        s.comb_hi = c.hi; s.comb_lo = c.lo; s.depth = mdepth;
    }
}

// ------------------------------------------------ tiles on the diagonal -----
// grid = (2, row tiles): CTA x handles the x-th column tile overlapping its row tile
// [r0,r1).  Columns left of r0 are a low side against the rows (split point between
// indices r0-1 and r0), columns from r1 on a high side (split between r1-1 and r1):
// both use the rows' precomputed prow/srow tables and two in-kernel queries per
// column.  Only the <= TR x TR block on the diagonal itself takes per-element
// queries, spread over all threads of the CTA.
__global__ void __launch_bounds__(MT)
k_matrix_diag(const TreeView tv, const int32_t *__restrict__ ids, const MatTables mt, int64_t n,
              int64_t row_begin, int64_t row_end, double *__restrict__ out) {
    __shared__ double s_hch[TR], s_hcl[TR], s_lch[TR], s_lcl[TR], s_ph[TR], s_pl[TR];
    __shared__ uint32_t s_hdep[TR], s_ldep[TR];
    __shared__ int32_t s_id[TR];
    __shared__ uint64_t s_suf[TR], s_pre[TR];
    __shared__ MidRec s_mid[2];

    const int64_t r0 = row_begin + int64_t(blockIdx.y) * TR;
    const int rows = int(row_end - r0 < TR ? row_end - r0 : TR);
    if (rows <= 0) return;
    const int64_t r1 = r0 + rows;
    const int64_t c0 = (r0 / TC + blockIdx.x) * TC;
    if (c0 >= n || c0 >= r1) return;
    const int64_t c1 = c0 + TC < n ? c0 + TC : n;
    const int t = threadIdx.x;

    if (t < 2) {  // middle ranges of the two split points
        MidRec m{0.0, 0.0, 0u, 0u, 0ull};
        const int64_t lowmax = t == 0 ? r0 - 1 : r1 - 1;
        if (lowmax >= 0 && lowmax + 1 < n) {
            const uint64_t k = mat_query(tv, mat_id(ids, lowmax), mat_id(ids, lowmax + 1));
            const dd r = st_ld_rd(tv, st_key_id(k));
            m = MidRec{r.hi, r.lo, uint32_t(k >> 32), 1u, 0ull};
        }
        s_mid[t] = m;
    }
    __syncthreads();
    if (t < rows) {
        const int64_t k = r0 + t;
        const int32_t id = mat_id(ids, k);
        const RecRaw r = st_ld_rec(tv, id);
        s_id[t] = id; s_suf[t] = r.suf; s_pre[t] = r.pre;
        s_ph[t] = r.rd_hi; s_pl[t] = r.rd_lo;
        const SideRec h = ld_side(mt.prow + (k - row_begin));  // rows as the high side
        s_hdep[t] = h.depth; s_hch[t] = h.comb_hi; s_hcl[t] = h.comb_lo;
        SideRec l = ld_side(mt.srow + (k - row_begin));        // rows as the low side
        const MidRec m = s_mid[1];
        if (m.valid) mat_apply_mid(l, dd{r.rd_hi, r.rd_lo}, m.depth, dd{m.rd_hi, m.rd_lo});
        s_ldep[t] = l.depth; s_lch[t] = l.comb_hi; s_lcl[t] = l.comb_lo;
    }
    // this thread's two columns: 0 = on the diagonal block, 1 = low side (left), 2 = high side (right)
    const int64_t cA = c0 + 2 * t;
    SideRec sd[2];
    dd pl[2];
    int role[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t c = cA + h;
        role[h] = 0;
        sd[h] = SideRec{0.0, 0.0, 0u, 0u, 0ull};
        pl[h] = dd{0.0, 0.0};
        if (c >= c1) { role[h] = -1; continue; }
        if (c >= r0 && c < r1) continue;
        const int32_t id = mat_id(ids, c);
        pl[h] = st_ld_rd(tv, id);
        if (c < r0) {
            role[h] = 1;
            sd[h] = mat_side(tv, pl[h], id, mat_id(ids, r0 - 1));
            const MidRec m = s_mid[0];
            mat_apply_mid(sd[h], pl[h], m.depth, dd{m.rd_hi, m.rd_lo});
        } else {
            role[h] = 2;
            sd[h] = mat_side(tv, pl[h], mat_id(ids, r1), id);
        }
    }
    __syncthreads();
    double *orow = out + (r0 - row_begin) * n;
    for (int i = 0; i < rows; ++i) {
        const dd rp{s_ph[i], s_pl[i]};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (role[h] <= 0) continue;
            double v;
            if (role[h] == 1) {  // column low, row high: low combined iff coldepth <= rowdepth
                v = sd[h].depth <= s_hdep[i] ? dd_add_to_double(dd{sd[h].comb_hi, sd[h].comb_lo}, rp)
                                             : dd_add_to_double(pl[h], dd{s_hch[i], s_hcl[i]});
            } else {             // row low, column high
                v = s_ldep[i] <= sd[h].depth ? dd_add_to_double(dd{s_lch[i], s_lcl[i]}, pl[h])
                                             : dd_add_to_double(rp, dd{sd[h].comb_hi, sd[h].comb_lo});
            }
            st_st_stream_f64(orow + int64_t(i) * n + cA + h, v);
        }
    }
    // the block on the diagonal: rows x [d0, d1), per-element queries, all threads
    const int64_t d0 = r0 > c0 ? r0 : c0, d1 = r1 < c1 ? r1 : c1;
    const int dcols = int(d1 - d0);
    if (dcols <= 0) return;
    const SmemTables g = st_global_tables(tv);
    for (int e = t; e < rows * dcols; e += MT) {
        const int i = e / dcols, j = e % dcols;
        const int jr = int(d0 - r0) + j;  // the column's index among this tile's rows
        const int32_t rid = s_id[i], cid = s_id[jr];
        double v = 0.0;
        if (rid != cid) {
            bool ft;
            const bool row_lo = rid < cid;
            const int a = row_lo ? i : jr, b = row_lo ? jr : i;  // a: low id, b: high id
            const uint64_t key = st_rmq(tv, g, s_id[a], s_id[b], s_suf[a], s_pre[b], &ft);
            const dd rm = st_mrca_rd(tv, g, key, ft);
            v = st_patristic(dd{s_ph[a], s_pl[a]}, dd{s_ph[b], s_pl[b]}, rm);
        }
        st_st_stream_f64(orow + int64_t(i) * n + d0 + j, v);
    }
}

// ---------------------------------------------------- generic (slow) tile ---
// per-element RMQ + one rd gather: diagonal tiles and unsorted id lists
__global__ void __launch_bounds__(MT)
k_matrix_generic(const TreeView tv, const int32_t *__restrict__ ids, int64_t n, int64_t row_begin,
                 int64_t row_end, double *__restrict__ out) {
    const int64_t r0 = row_begin + int64_t(blockIdx.y) * TR;
    const int rows = int(row_end - r0 < TR ? row_end - r0 : TR);
    if (rows <= 0) return;
    const int64_t c0 = int64_t(blockIdx.x) * TC;
    if (c0 >= n) return;
    const int64_t c1 = c0 + TC < n ? c0 + TC : n;
    __shared__ int32_t s_id[TR];
    __shared__ double s_ph[TR], s_pl[TR];
    __shared__ uint64_t s_suf[TR], s_pre[TR];
    const int t = threadIdx.x;
    const int64_t cA = c0 + 2 * t, cB = cA + 1;
    const bool hasA = cA < c1, hasB = cB < c1;
    const int32_t idA = hasA ? mat_id(ids, cA) : 0, idB = hasB ? mat_id(ids, cB) : 0;
    for (int i = t; i < rows; i += MT) {
        int32_t id = mat_id(ids, r0 + i);
        RecRaw r = st_ld_rec(tv, id);
        s_id[i] = id;
        s_ph[i] = r.rd_hi; s_pl[i] = r.rd_lo;
        s_suf[i] = r.suf;  s_pre[i] = r.pre;
    }
    RecRaw ra{}, rb{};
    if (hasA) ra = st_ld_rec(tv, idA);
    if (hasB) rb = st_ld_rec(tv, idB);
    __syncthreads();
    const SmemTables g = st_global_tables(tv);
    for (int i = 0; i < rows; ++i) {
        const int32_t rid = s_id[i];
        const dd rrd{s_ph[i], s_pl[i]};
        double v[2] = {0.0, 0.0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool has = h ? hasB : hasA;
            const int32_t cid = h ? idB : idA;
            const RecRaw &rc = h ? rb : ra;
            if (has && cid != rid) {
                bool ft;
                uint64_t key = rid < cid ? st_rmq(tv, g, rid, cid, s_suf[i], rc.pre, &ft)
                                         : st_rmq(tv, g, cid, rid, rc.suf, s_pre[i], &ft);
                dd rm = st_mrca_rd(tv, g, key, ft);
                // low-id operand first, as in the pair kernel
                v[h] = rid < cid ? st_patristic(rrd, dd{rc.rd_hi, rc.rd_lo}, rm)
                                 : st_patristic(dd{rc.rd_hi, rc.rd_lo}, rrd, rm);
            }
        }
        mat_store2(out + (r0 + i - row_begin) * n, cA, n, v[0], v[1]);
    }
}

// ------------------------------------------------------------- tile kernel --
// Ordered tiles read their side data from the set-up tables: no index query, no
// gather.  Every element is one depth compare and
//     d = lowdepth <= highdepth ? lowcomb + highplain : lowplain + highcomb
// in double-double (9 fp64 adds) -- or, when every low word of the tile's operands
// is zero (always the case for trees whose root distances are exact in fp64), in
// ONE fp64 add with a bit-identical result.
__global__ void __launch_bounds__(MT)
k_matrix_ordered(const TreeView tv, const int32_t *__restrict__ ids, const MatTables mt, int64_t n,
                 int64_t row_begin, int64_t row_end, double *__restrict__ out) {
    __shared__ double s_ch[TR], s_cl[TR], s_ph[TR], s_pl[TR];
    __shared__ uint32_t s_dep[TR];
    __shared__ __align__(32) SideRec s_colside[TC];  // 16 KB
    __shared__ __align__(16) double2 s_colrd[TC];    //  8 KB
    __shared__ __align__(8) uint64_t s_bar;

    const int64_t r0 = row_begin + int64_t(blockIdx.y) * TR;
    const int64_t c0 = int64_t(blockIdx.x) * TC;
    const int rows = int(row_end - r0 < TR ? row_end - r0 : TR);
    const int cols = int(n - c0 < TC ? n - c0 : TC);
    if (rows <= 0 || cols <= 0) return;
    const int64_t r1 = r0 + rows, c1 = c0 + cols;  // exclusive
    const int t = threadIdx.x;

    // rows are the low side (above the diagonal), or the columns are; tiles that
    // overlap the diagonal belong to k_matrix_diag
    if (!(r1 <= c0) && !(c1 <= r0)) return;
    const bool rows_low = r1 <= c0;
    const MidRec mid = mt.mid[int64_t(blockIdx.y) * mt.n_ct + blockIdx.x];
    const dd midrd{mid.rd_hi, mid.rd_lo};

    int nonzero_lo = 0;
    if (t < rows) {
        const int64_t k = r0 + t;
        const double2 p = __ldg(mt.rd + k);
        SideRec s = ld_side((rows_low ? mt.srow : mt.prow) + (k - row_begin));
        if (rows_low) mat_apply_mid(s, dd{p.x, p.y}, mid.depth, midrd);
        s_dep[t] = s.depth;
        s_ch[t] = s.comb_hi; s_cl[t] = s.comb_lo;
        s_ph[t] = p.x;       s_pl[t] = p.y;
        nonzero_lo |= (s.comb_lo != 0.0) | (p.y != 0.0);
    }
    // the tile's column block (side records + root distances: two contiguous runs of the
    // set-up tables) is staged into shared memory by TMA bulk copies, completion on an mbarrier
    {
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (t == 0) {
            const uint32_t bytes_side = uint32_t(cols) * 32u, bytes_rd = uint32_t(cols) * 16u;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes_side + bytes_rd)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s_colside)),
                         "l"((rows_low ? mt.pcol : mt.scol) + c0), "r"(bytes_side), "r"(bar_a)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s_colrd)),
                         "l"(mt.rd + c0), "r"(bytes_rd), "r"(bar_a)
                         : "memory");
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar_a), "r"(0)
                : "memory");
        }
    }
    // this thread's two columns
    const int64_t cA = c0 + 2 * t, cB = cA + 1;
    const bool hasA = cA < c1, hasB = cB < c1;
    SideRec a{}, b{};
    dd pa{0.0, 0.0}, pb{0.0, 0.0};
    if (hasA) {
        const double2 p = s_colrd[2 * t];
        pa = dd{p.x, p.y};
        a = s_colside[2 * t];
        if (!rows_low) mat_apply_mid(a, pa, mid.depth, midrd);
        nonzero_lo |= (a.comb_lo != 0.0) | (pa.lo != 0.0);
    }
    if (hasB) {
        const double2 p = s_colrd[2 * t + 1];
        pb = dd{p.x, p.y};
        b = s_colside[2 * t + 1];
        if (!rows_low) mat_apply_mid(b, pb, mid.depth, midrd);
        nonzero_lo |= (b.comb_lo != 0.0) | (pb.lo != 0.0);
    }
    const int slow = __syncthreads_or(nonzero_lo);
    double *orow = out + (r0 - row_begin) * n;

    if (!slow) {
        // "low side combined" iff lowdepth <= highdepth; rows_low: row is the low side
#pragma unroll 4
        for (int i = 0; i < rows; ++i) {
            const uint32_t rdp = s_dep[i];
            const double rc = s_ch[i], rp = s_ph[i];
            const bool selA = rows_low ? (rdp <= a.depth) : (rdp < a.depth);
            const bool selB = rows_low ? (rdp <= b.depth) : (rdp < b.depth);
            // sel true -> row comb + column plain ; false -> row plain + column comb
            const double v0 = selA ? __dadd_rn(rc, pa.hi) : __dadd_rn(rp, a.comb_hi);
            const double v1 = selB ? __dadd_rn(rc, pb.hi) : __dadd_rn(rp, b.comb_hi);
            mat_store2(orow + int64_t(i) * n, cA, n, v0, v1);
        }
        return;
    }
    for (int i = 0; i < rows; ++i) {
        const uint32_t rdp = s_dep[i];
        const dd rc{s_ch[i], s_cl[i]}, rp{s_ph[i], s_pl[i]};
        const bool selA = rows_low ? (rdp <= a.depth) : (rdp < a.depth);
        const bool selB = rows_low ? (rdp <= b.depth) : (rdp < b.depth);
        const double v0 = selA ? dd_add_to_double(rc, pa) : dd_add_to_double(rp, dd{a.comb_hi, a.comb_lo});
        const double v1 = selB ? dd_add_to_double(rc, pb) : dd_add_to_double(rp, dd{b.comb_hi, b.comb_lo});
        mat_store2(orow + int64_t(i) * n, cA, n, v0, v1);
    }
}

__global__ void k_narrow_ids(int64_t n, const int64_t *__restrict__ in, int32_t *__restrict__ out,
                             int32_t n_nodes, RangeStatus *status, int *unsorted) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long v = in[i];
    if ((unsigned long long)v >= (unsigned long long)n_nodes) {
        if (v >= n_nodes) atomicMax(&status->max_bad, (unsigned long long)v);
        else atomicMin(&status->min_bad, v);
        v = 0;
    }
    out[i] = int32_t(v);
    if (i + 1 < n && in[i + 1] <= in[i]) *unsorted = 1;
}

// Set-up tables + tile kernel on one stream.  The tables live in stream-ordered
// allocations (cudaMallocAsync), so the call stays asynchronous.
static int launch_matrix(const st_tree *t, const int32_t *d_ids, bool sorted, int64_t n,
                         int64_t row_begin, int64_t row_end, double *d_out, cudaStream_t s) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0 || n <= 0) return ST_OK;
    dim3 grid((unsigned)((n + TC - 1) / TC), (unsigned)((rows + TR - 1) / TR));
    if (grid.y > 65535u) {
        st_set_error("st_distance_matrix: more than 65535*%d rows per call", TR);
        return ST_ERR_INVALID_ARG;
    }
    MatTables mt{};
    if (!sorted) {
        k_matrix_generic<<<grid, MT, 0, s>>>(t->view, d_ids, n, row_begin, row_end, d_out);
        ST_CUDA(cudaGetLastError());
        return ST_OK;
    }
    const int64_t n_rt = grid.y, n_ct = grid.x;
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };  // records are read with 32-byte loads
    const size_t bytes_rd = pad(size_t(n) * sizeof(double2)), bytes_col = pad(size_t(n) * sizeof(SideRec));
    const size_t bytes_row = pad(size_t(rows) * sizeof(SideRec)), bytes_mid = pad(size_t(n_rt) * n_ct * sizeof(MidRec));
    const size_t total = bytes_rd + 2 * bytes_col + 2 * bytes_row + bytes_mid;
    unsigned char *base = nullptr;
    ST_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&base), total, s));
    unsigned char *p = base;
    double2 *rd = reinterpret_cast<double2 *>(p); p += bytes_rd;
    SideRec *pcol = reinterpret_cast<SideRec *>(p); p += bytes_col;
    SideRec *scol = reinterpret_cast<SideRec *>(p); p += bytes_col;
    SideRec *prow = reinterpret_cast<SideRec *>(p); p += bytes_row;
    SideRec *srow = reinterpret_cast<SideRec *>(p); p += bytes_row;
    MidRec *mid = reinterpret_cast<MidRec *>(p);
    k_matrix_sides<<<unsigned((n + 255) / 256), 256, 0, s>>>(t->view, d_ids, n, row_begin, row_end, rd, pcol,
                                                            scol, prow, srow);
    k_matrix_mid<<<unsigned((n_rt * n_ct + 255) / 256), 256, 0, s>>>(t->view, d_ids, n, row_begin, row_end,
                                                                    int32_t(n_rt), int32_t(n_ct), mid);
    mt = MatTables{rd, pcol, scol, prow, srow, mid, int32_t(n_ct)};
    k_matrix_ordered<<<grid, MT, 0, s>>>(t->view, d_ids, mt, n, row_begin, row_end, d_out);
    k_matrix_diag<<<dim3(2, grid.y), MT, 0, s>>>(t->view, d_ids, mt, n, row_begin, row_end, d_out);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(base, s);
    if (e != cudaSuccess) {
        st_set_error("st_distance_matrix: launch failed: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}

extern "C" int st_distance_matrix(const st_tree *t, const int64_t *ids, int64_t n, int64_t row_begin,
                                  int64_t row_end, double *out, int out_on_device, void *stream) {
    if (!t || n < 0 || row_begin < 0 || row_end > n || row_begin > row_end || (!out && row_end > row_begin)) {
        st_set_error("st_distance_matrix: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    if (!ids && n != t->n_leaves) {
        st_set_error("st_distance_matrix: ids == NULL means all %lld leaves, got n = %lld",
                     (long long)t->n_leaves, (long long)n);
        return ST_ERR_INVALID_ARG;
    }
    if (n == 0 || row_begin == row_end) return ST_OK;
    DeviceGuard g(t->device);
    if (out_on_device && !ids)  // nothing of the host is involved: asynchronous on the caller's stream
        return launch_matrix(t, nullptr, true, n, row_begin, row_end, out, static_cast<cudaStream_t>(stream));

    // a host id list and/or a host result: this call's streams, scratch and status word
    // come from a lane of the device's host context
    LaneGuard lg(t->device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    cudaStream_t s = out_on_device ? static_cast<cudaStream_t>(stream) : lane->streams[0];

    // node list -> device int32 (+ range check, + sortedness)
    int32_t *d_ids = nullptr;
    bool sorted = true;
    int rc = ST_OK;
    if (ids) {
        int64_t *d_ids64 = nullptr;
        int *d_flag = nullptr;
        // stream-ordered scratch (recycled by the pool: a cudaMalloc / cudaFree pair costs more than a
        // small matrix does, and cudaFree synchronises the whole device)
        if (cudaMallocAsync(reinterpret_cast<void **>(&d_ids), size_t(n) * 4, s) != cudaSuccess ||
            cudaMallocAsync(reinterpret_cast<void **>(&d_ids64), size_t(n) * 8, s) != cudaSuccess ||
            cudaMallocAsync(reinterpret_cast<void **>(&d_flag), 4, s) != cudaSuccess) {
            cudaGetLastError();
            if (d_ids) cudaFreeAsync(d_ids, s);
            if (d_ids64) cudaFreeAsync(d_ids64, s);
            cudaStreamSynchronize(s);
            st_set_error("st_distance_matrix: device allocation failed");
            return ST_ERR_NOMEM;
        }
        cudaMemcpyAsync(d_ids64, ids, size_t(n) * 8, cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_flag, 0, 4, s);
        k_narrow_ids<<<unsigned((n + 255) / 256), 256, 0, s>>>(n, d_ids64, d_ids, int32_t(t->n_nodes),
                                                              lane->d_status, d_flag);
        int unsorted = 0;
        cudaMemcpyAsync(&unsorted, d_flag, 4, cudaMemcpyDeviceToHost, s);
        unsigned long long mxb = 0;
        long long mnb = 0;
        rc = st_lane_read_status(lane, s, &mxb, &mnb);  // synchronises s
        cudaFreeAsync(d_ids64, s);
        cudaFreeAsync(d_flag, s);
        if (rc == ST_OK && (mxb != 0 || mnb != 0)) {
            st_set_bad_node(mxb != 0 ? (int64_t)mxb : (int64_t)mnb);
            st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(), (long long)t->n_nodes);
            rc = ST_ERR_NODE_RANGE;
        }
        if (rc != ST_OK) {
            cudaFreeAsync(d_ids, s);
            return rc;
        }
        sorted = !unsorted;
    }

    if (out_on_device) {
        rc = launch_matrix(t, d_ids, sorted, n, row_begin, row_end, out, s);
        if (d_ids) cudaFreeAsync(d_ids, s);  // ordered after the kernels on s
        cudaStreamSynchronize(s);
        return rc;
    }

    // host output: row bands through two device buffers on two of the lane's streams -- the
    // kernel of band k+1 runs while band k crosses PCIe.  A page-locked result (the Python
    // shim's arrays come from the pinned pool) takes the D2H copies directly at PCIe rate;
    // a pageable one goes through the driver's staged copy.
    const int64_t band_bytes = int64_t(64) << 20;
    int64_t band_rows = std::max<int64_t>(TR, (band_bytes / (n * 8)) / TR * TR);
    band_rows = std::min<int64_t>(band_rows, ((row_end - row_begin) + TR - 1) / TR * TR);
    double *d_band[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        if (cudaMallocAsync(reinterpret_cast<void **>(&d_band[i]), size_t(band_rows) * n * 8, lane->streams[i]) !=
            cudaSuccess) {
            cudaGetLastError();
            if (d_band[0]) cudaFreeAsync(d_band[0], lane->streams[0]);
            cudaStreamSynchronize(lane->streams[0]);
            if (d_ids) cudaFreeAsync(d_ids, s);
            st_set_error("st_distance_matrix: device allocation of a %lld-row band failed", (long long)band_rows);
            return ST_ERR_NOMEM;
        }
    }
    int k = 0;
    for (int64_t r = row_begin; r < row_end && rc == ST_OK; r += band_rows, ++k) {
        const int b = k & 1;
        cudaStream_t sb = lane->streams[b];
        const int64_t re = std::min(row_end, r + band_rows);
        rc = launch_matrix(t, d_ids, sorted, n, r, re, d_band[b], sb);
        if (rc != ST_OK) break;
        if (cudaMemcpyAsync(out + (r - row_begin) * n, d_band[b], size_t(re - r) * n * 8,
                            cudaMemcpyDeviceToHost, sb) != cudaSuccess) {
            st_set_error("st_distance_matrix: D2H copy failed");
            rc = ST_ERR_CUDA;
        }
    }
    for (int i = 0; i < 2; ++i) cudaFreeAsync(d_band[i], lane->streams[i]);
    cudaStreamSynchronize(lane->streams[0]);
    cudaStreamSynchronize(lane->streams[1]);
    if (rc == ST_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            st_set_error("st_distance_matrix: %s", cudaGetErrorString(e));
            rc = ST_ERR_CUDA;
        }
    }
    if (d_ids) cudaFreeAsync(d_ids, s);  // both streams have been synchronised: nothing reads it any more
    return rc;
}
