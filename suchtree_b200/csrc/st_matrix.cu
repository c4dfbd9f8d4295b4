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

static const int MT = 256;       // threads per CTA
static const int TC = 2 * MT;    // columns per tile (two per thread -> 16-byte stores)
static const int TR = 64;        // rows per tile

struct SideData {  // per row (shared memory) or per column (registers)
    uint64_t key;
    dd comb, plain;
};

__device__ __forceinline__ int32_t mat_id(const int32_t *__restrict__ ids, int64_t k) {
    return ids ? __ldg(ids + k) : int32_t(2 * k);  // default: leaf k has id 2k
}

// full query through global-memory tables (few per tile; L1/L2 cached)
__device__ __forceinline__ uint64_t mat_query(const TreeView &tv, int32_t a, int32_t b) {
    int32_t lo = min(a, b), hi = max(a, b);
    if (lo == hi) return st_key(__ldg(tv.depth + lo), lo);
    RecRaw rl = st_ld_rec(tv.rec + lo), rh = st_ld_rec(tv.rec + hi);
    SmemTables g{tv.stk, tv.brd};
    bool ft;
    return st_rmq(tv, g, lo, hi, rl.suf, rh.pre, &ft);
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

template <bool SORTED>
__global__ void __launch_bounds__(MT)
k_matrix(const TreeView tv, const int32_t *__restrict__ ids, int64_t n, int64_t row_begin,
         int64_t row_end, double *__restrict__ out) {
    __shared__ int32_t s_id[TR];
    __shared__ uint64_t s_key[TR];
    __shared__ double s_ch[TR], s_cl[TR], s_ph[TR], s_pl[TR];
    __shared__ uint64_t s_suf[TR], s_pre[TR];
    __shared__ uint64_t s_mid;
    __shared__ dd s_midrd;

    const int64_t r0 = row_begin + int64_t(blockIdx.y) * TR;
    const int64_t c0 = int64_t(blockIdx.x) * TC;
    const int rows = int(row_end - r0 < TR ? row_end - r0 : TR);
    const int cols = int(n - c0 < TC ? n - c0 : TC);
    if (rows <= 0 || cols <= 0) return;
    const int64_t r1 = r0 + rows, c1 = c0 + cols;  // exclusive
    const int t = threadIdx.x;

    // tile class: 0 = per-element (diagonal / unsorted), 1 = rows are the low side, 2 = columns are
    int cls = 0;
    if (SORTED) {
        if (r1 <= c0) cls = 1;
        else if (c1 <= r0) cls = 2;
    }

    // this thread's two columns
    const int64_t cA = c0 + 2 * t, cB = cA + 1;
    const bool hasA = cA < c1, hasB = cB < c1;
    const int32_t idA = hasA ? mat_id(ids, cA) : 0, idB = hasB ? mat_id(ids, cB) : 0;

    if (cls == 0) {
        // ---- generic tile: per-element RMQ + one rd gather
        for (int i = t; i < rows; i += MT) {
            int32_t id = mat_id(ids, r0 + i);
            RecRaw r = st_ld_rec(tv.rec + id);
            s_id[i] = id;
            s_ph[i] = r.rd_hi; s_pl[i] = r.rd_lo;
            s_suf[i] = r.suf;  s_pre[i] = r.pre;
        }
        RecRaw ra{}, rb{};
        if (hasA) ra = st_ld_rec(tv.rec + idA);
        if (hasB) rb = st_ld_rec(tv.rec + idB);
        __syncthreads();
        SmemTables g{tv.stk, tv.brd};
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
                    dd rm = st_ld_rd(tv.rec + st_key_id(key));
                    // low-id operand first, as in the pair kernel
                    v[h] = rid < cid ? st_patristic(rrd, dd{rc.rd_hi, rc.rd_lo}, rm)
                                     : st_patristic(dd{rc.rd_hi, rc.rd_lo}, rrd, rm);
                }
            }
            mat_store2(out + (r0 + i - row_begin) * n, cA, n, v[0], v[1]);
        }
        return;
    }

    // ---- ordered tile.  low side = rows (cls 1) or columns (cls 2)
    const bool rows_low = cls == 1;
    const int32_t lowmax = rows_low ? mat_id(ids, r1 - 1) : mat_id(ids, c1 - 1);
    const int32_t highmin = rows_low ? mat_id(ids, c0) : mat_id(ids, r0);
    if (t == 0) {
        uint64_t k = mat_query(tv, lowmax, highmin);
        s_mid = k;
        s_midrd = st_ld_rd(tv.rec + st_key_id(k));
    }
    __syncthreads();
    const uint64_t mid = s_mid;
    const dd midrd = s_midrd;

    auto side = [&](int32_t id, bool low) {
        SideData s;
        dd rd = st_ld_rd(tv.rec + id);
        s.plain = rd;
        if (low) {
            uint64_t k = mat_query(tv, id, lowmax);
            if (k <= mid) {
                s.key = k;
                s.comb = dd_minus_2x(rd, st_ld_rd(tv.rec + st_key_id(k)));
            } else {
                s.key = mid;
                s.comb = dd_minus_2x(rd, midrd);
            }
        } else {
            uint64_t k = mat_query(tv, highmin, id);
            s.key = k;
            s.comb = dd_minus_2x(rd, st_ld_rd(tv.rec + st_key_id(k)));
        }
        return s;
    };

    for (int i = t; i < rows; i += MT) {
        SideData s = side(mat_id(ids, r0 + i), rows_low);
        s_key[i] = s.key;
        s_ch[i] = s.comb.hi;  s_cl[i] = s.comb.lo;
        s_ph[i] = s.plain.hi; s_pl[i] = s.plain.lo;
    }
    SideData a{}, b{};
    if (hasA) a = side(idA, !rows_low);
    if (hasB) b = side(idB, !rows_low);
    __syncthreads();

    for (int i = 0; i < rows; ++i) {
        const uint64_t rk = s_key[i];
        const dd rc{s_ch[i], s_cl[i]}, rp{s_ph[i], s_pl[i]};
        // "low side combined" iff lowkey <= highkey
        const bool selA = rows_low ? (rk <= a.key) : !(a.key <= rk);
        const bool selB = rows_low ? (rk <= b.key) : !(b.key <= rk);
        // selX true -> row comb + column plain ; false -> row plain + column comb
        double v0 = selA ? dd_add_to_double(rc, a.plain) : dd_add_to_double(rp, a.comb);
        double v1 = selB ? dd_add_to_double(rc, b.plain) : dd_add_to_double(rp, b.comb);
        mat_store2(out + (r0 + i - row_begin) * n, cA, n, v0, v1);
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

static int launch_matrix(const st_tree *t, const int32_t *d_ids, bool sorted, int64_t n,
                         int64_t row_begin, int64_t row_end, double *d_out, cudaStream_t s) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0 || n <= 0) return ST_OK;
    dim3 grid((unsigned)((n + TC - 1) / TC), (unsigned)((rows + TR - 1) / TR));
    if (grid.y > 65535u) {
        st_set_error("st_distance_matrix: more than 65535*%d rows per call", TR);
        return ST_ERR_INVALID_ARG;
    }
    if (sorted)
        k_matrix<true><<<grid, MT, 0, s>>>(t->view, d_ids, n, row_begin, row_end, d_out);
    else
        k_matrix<false><<<grid, MT, 0, s>>>(t->view, d_ids, n, row_begin, row_end, d_out);
    ST_CUDA(cudaGetLastError());
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
    cudaStream_t s = out_on_device ? static_cast<cudaStream_t>(stream) : t->streams[0];

    // node list -> device int32 (+ range check, + sortedness)
    int32_t *d_ids = nullptr;
    bool sorted = true;
    int64_t *d_ids64 = nullptr;
    int *d_flag = nullptr;
    int rc = ST_OK;
    auto cleanup = [&]() {
        cudaFree(d_ids);
        cudaFree(d_ids64);
        cudaFree(d_flag);
    };
    if (ids) {
        ST_CUDA(cudaMalloc(&d_ids, size_t(n) * 4));
        if (cudaMalloc(&d_ids64, size_t(n) * 8) != cudaSuccess || cudaMalloc(&d_flag, 4) != cudaSuccess) {
            cleanup();
            st_set_error("st_distance_matrix: cudaMalloc failed");
            return ST_ERR_NOMEM;
        }
        cudaMemcpyAsync(d_ids64, ids, size_t(n) * 8, cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_flag, 0, 4, s);
        k_narrow_ids<<<unsigned((n + 255) / 256), 256, 0, s>>>(n, d_ids64, d_ids, int32_t(t->n_nodes),
                                                              t->d_status, d_flag);
        int unsorted = 0;
        cudaMemcpyAsync(&unsorted, d_flag, 4, cudaMemcpyDeviceToHost, s);
        bool bad = false;
        rc = st_read_range_status(t, s, &bad);  // synchronises s
        if (rc == ST_OK && bad) rc = ST_ERR_NODE_RANGE;
        if (rc != ST_OK) {
            cleanup();
            return rc;
        }
        sorted = !unsorted;
    }

    if (out_on_device) {
        rc = launch_matrix(t, d_ids, sorted, n, row_begin, row_end, out, s);
        if (d_ids) cudaStreamSynchronize(s);  // d_ids is freed below
        cleanup();
        return rc;
    }

    // host output: row bands through two device buffers / two streams
    std::lock_guard<std::mutex> lock(t->host_mu);
    const int64_t band_bytes = int64_t(256) << 20;
    int64_t band_rows = std::max<int64_t>(TR, (band_bytes / (n * 8)) / TR * TR);
    band_rows = std::min<int64_t>(band_rows, ((row_end - row_begin) + TR - 1) / TR * TR);
    double *d_band[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        if (cudaMalloc(&d_band[i], size_t(band_rows) * n * 8) != cudaSuccess) {
            cudaFree(d_band[0]);
            cleanup();
            st_set_error("st_distance_matrix: cudaMalloc of a %lld-row band failed", (long long)band_rows);
            return ST_ERR_NOMEM;
        }
    }
    int k = 0;
    for (int64_t r = row_begin; r < row_end && rc == ST_OK; r += band_rows, ++k) {
        const int b = k & 1;
        cudaStream_t sb = t->streams[b];
        const int64_t re = std::min(row_end, r + band_rows);
        rc = launch_matrix(t, d_ids, sorted, n, r, re, d_band[b], sb);
        if (rc != ST_OK) break;
        if (cudaMemcpyAsync(out + (r - row_begin) * n, d_band[b], size_t(re - r) * n * 8,
                            cudaMemcpyDeviceToHost, sb) != cudaSuccess) {
            st_set_error("st_distance_matrix: D2H copy failed");
            rc = ST_ERR_CUDA;
        }
    }
    cudaStreamSynchronize(t->streams[0]);
    cudaStreamSynchronize(t->streams[1]);
    if (rc == ST_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            st_set_error("st_distance_matrix: %s", cudaGetErrorString(e));
            rc = ST_ERR_CUDA;
        }
    }
    cudaFree(d_band[0]);
    cudaFree(d_band[1]);
    cleanup();
    return rc;
}
