// Batched quartet topologies.
//
// Replaces SuchTree._quartet_topologies (MuchTree.pyx:1331-1376): six _mrca calls
// per quartet (each an O(depth^2) pointer chase in the reference) become six
// range-minimum lookups over the four endpoint records -- 4 random sectors (2 in
// the compact layout's half-sector records) + 64 B streamed (int64x4 in, int64x4
// out) per quartet.  The selection rule is the reference's, literally: C[j] = how
// many of the six MRCAs equal M[j]; the first j with C[j] == 1 picks row j of the
// permutation table I (:1319-1320); no such j -> row 5.
#include <algorithm>

#include "st_device.cuh"
#include "st_hostctx.cuh"

static const int QQT = 256;

struct Quad {
    long long v[4];
};

__device__ __forceinline__ Quad quad_load(const int64_t *p, bool aligned) {
    Quad q;
    if (aligned) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(q.v[0]), "=l"(q.v[1]), "=l"(q.v[2]), "=l"(q.v[3])
                     : "l"(p));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) q.v[k] = __ldg(reinterpret_cast<const long long *>(p) + k);
    }
    return q;
}
__device__ __forceinline__ void quad_store(int64_t *p, bool aligned, long long a, long long b,
                                           long long c, long long d) {
    if (aligned) {
        asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b),
                     "l"(c), "l"(d)
                     : "memory");
    } else {
        p[0] = a; p[1] = b; p[2] = c; p[3] = d;
    }
}

template <int M>
__global__ void __launch_bounds__(QQT)
k_quartets(const TreeView tv, const int64_t *__restrict__ quartets, int64_t n,
           int64_t *__restrict__ out, int aligned) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tables_bar;
    const SmemTables sm = st_load_tables<M>(tv, smem_raw, &tables_bar);
    const long long nn = tv.n_nodes;
    for (int64_t i = int64_t(blockIdx.x) * QQT + threadIdx.x; i < n; i += int64_t(gridDim.x) * QQT) {
        const Quad q = quad_load(quartets + 4 * i, aligned != 0);
        // range check of quartet_topologies_bulk (MuchTree.pyx:1303-1310), on the device
        long long mx = q.v[0], mn = q.v[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            mx = q.v[k] > mx ? q.v[k] : mx;
            mn = q.v[k] < mn ? q.v[k] : mn;
        }
        if (mn < 0 || mx >= nn) {
            if (mx >= nn) atomicMax(&tv.status->max_bad, (unsigned long long)mx);
            if (mn < 0) atomicMin(&tv.status->min_bad, mn);
            quad_store(out + 4 * i, aligned != 0, -1, -1, -1, -1);
            continue;
        }
        int32_t id[4];
        RecRaw r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = int32_t(q.v[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = st_ld_rec<M>(tv, id[k]);
        // the six pairs in the reference's order: ab ac ad bc bd cd
        int32_t m[6];
        int p = 0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
#pragma unroll
            for (int y = x + 1; y < 4; ++y, ++p) {
                if (id[x] == id[y]) {
                    m[p] = id[x];
                } else {
                    const bool xl = id[x] < id[y];
                    bool ft;
                    const uint64_t k = st_rmq<M>(tv, sm, xl ? id[x] : id[y], xl ? id[y] : id[x],
                                                 xl ? r[x].suf : r[y].suf, xl ? r[y].pre : r[x].pre, &ft);
                    m[p] = st_mrca_id<M>(tv, sm, k, ft);
                }
            }
        }
        int j = 5;
#pragma unroll
        for (int a = 5; a >= 0; --a) {
            int c = 0;
#pragma unroll
            for (int b = 0; b < 6; ++b) c += (m[a] == m[b]);
            if (c == 1) j = a;  // descending scan: the smallest such index wins
        }
        // rows of I: {0,1,2,3} {0,2,1,3} {0,3,1,2} {1,2,0,3} {1,3,0,2} {2,3,0,1}
        long long o0, o1, o2, o3;
        switch (j) {
            case 0: o0 = q.v[0]; o1 = q.v[1]; o2 = q.v[2]; o3 = q.v[3]; break;
            case 1: o0 = q.v[0]; o1 = q.v[2]; o2 = q.v[1]; o3 = q.v[3]; break;
            case 2: o0 = q.v[0]; o1 = q.v[3]; o2 = q.v[1]; o3 = q.v[2]; break;
            case 3: o0 = q.v[1]; o1 = q.v[2]; o2 = q.v[0]; o3 = q.v[3]; break;
            case 4: o0 = q.v[1]; o1 = q.v[3]; o2 = q.v[0]; o3 = q.v[2]; break;
            default: o0 = q.v[2]; o1 = q.v[3]; o2 = q.v[0]; o3 = q.v[1]; break;
        }
        quad_store(out + 4 * i, aligned != 0, o0, o1, o2, o3);
    }
}

template <int M>
static int launch_quartets_m(const st_tree *t, const int64_t *d_q, int64_t n, int64_t *d_out,
                             cudaStream_t stream, RangeStatus *status) {
    auto kern = k_quartets<M>;
    const int smem = t->query_smem_bytes;
    int rc = st_raise_smem(kern, t->device, smem);
    if (rc != ST_OK) return rc;
    int per_sm = 0;
    ST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QQT, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = int(std::min<int64_t>((n + QQT - 1) / QQT, int64_t(t->sm_count) * per_sm));
    const int aligned = (reinterpret_cast<uintptr_t>(d_q) % 32 == 0) && (reinterpret_cast<uintptr_t>(d_out) % 32 == 0);
    TreeView view = t->view;
    if (status) view.status = status;
    kern<<<grid, QQT, smem, stream>>>(view, d_q, n, d_out, aligned);
    ST_CUDA(cudaGetLastError());
    return ST_OK;
}

static int launch_quartets(const st_tree *t, const int64_t *d_q, int64_t n, int64_t *d_out,
                           cudaStream_t stream, RangeStatus *status = nullptr) {
    if (n == 0) return ST_OK;
    if (t->compact) return launch_quartets_m<1>(t, d_q, n, d_out, stream, status);
    if (t->compact_tables) return launch_quartets_m<3>(t, d_q, n, d_out, stream, status);
    return launch_quartets_m<0>(t, d_q, n, d_out, stream, status);
}

extern "C" int st_quartet_topologies_device(const st_tree *t, const int64_t *d_quartets, int64_t n,
                                            int64_t *d_out, void *stream) {
    if (!t || n < 0 || (n > 0 && (!d_quartets || !d_out))) {
        st_set_error("st_quartet_topologies_device: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    DeviceGuard g(t->device);
    return launch_quartets(t, d_quartets, n, d_out, static_cast<cudaStream_t>(stream));
}

// host buffers: chunks through stream-ordered device scratch, H2D | kernel | D2H
// overlapped across two streams
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
    const bool contiguous = (s0 == 4 && s1 == 1);
    const int64_t C = std::min<int64_t>(n, int64_t(1) << 21);
    int64_t *d_in[2] = {nullptr, nullptr}, *d_o[2] = {nullptr, nullptr};
    int64_t *h_pack = nullptr;
    auto cleanup = [&]() {
        for (int i = 0; i < 2; ++i) {
            if (d_in[i]) cudaFreeAsync(d_in[i], lane->streams[i]);
            if (d_o[i]) cudaFreeAsync(d_o[i], lane->streams[i]);
        }
        if (h_pack) cudaFreeHost(h_pack);
    };
    for (int i = 0; i < 2; ++i) {
        if (cudaMallocAsync(reinterpret_cast<void **>(&d_in[i]), size_t(C) * 32, lane->streams[i]) != cudaSuccess ||
            cudaMallocAsync(reinterpret_cast<void **>(&d_o[i]), size_t(C) * 32, lane->streams[i]) != cudaSuccess) {
            cleanup();
            st_set_error("st_quartet_topologies: device allocation failed");
            return ST_ERR_NOMEM;
        }
    }
    if (!contiguous && cudaMallocHost(&h_pack, size_t(C) * 32 * 2) != cudaSuccess) {
        cleanup();
        st_set_error("st_quartet_topologies: pinned allocation failed");
        return ST_ERR_NOMEM;
    }
    int rc = ST_OK;
    int c = 0;
    for (int64_t done = 0; done < n && rc == ST_OK; ++c) {
        const int b = c & 1;
        cudaStream_t st = lane->streams[b];
        const int64_t m = std::min(C, n - done);
        const int64_t *src = quartets + done * s0;
        if (!contiguous) {
            if (c >= 2) cudaStreamSynchronize(st);  // the pack buffer of this slot is free again
            int64_t *hp = h_pack + size_t(b) * C * 4;
            for (int64_t i = 0; i < m; ++i)
                for (int k = 0; k < 4; ++k) hp[4 * i + k] = src[i * s0 + k * s1];
            src = hp;
        }
        if (cudaMemcpyAsync(d_in[b], src, size_t(m) * 32, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            st_set_error("st_quartet_topologies: H2D copy failed");
            rc = ST_ERR_CUDA;
            break;
        }
        rc = launch_quartets(t, d_in[b], m, d_o[b], st, lane->d_status);
        if (rc != ST_OK) break;
        if (cudaMemcpyAsync(out + done * 4, d_o[b], size_t(m) * 32, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            st_set_error("st_quartet_topologies: D2H copy failed");
            rc = ST_ERR_CUDA;
        }
        done += m;
    }
    cudaError_t e0 = cudaStreamSynchronize(lane->streams[0]), e1 = cudaStreamSynchronize(lane->streams[1]);
    cleanup();
    if (rc == ST_OK && (e0 != cudaSuccess || e1 != cudaSuccess)) {
        st_set_error("st_quartet_topologies: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
        rc = ST_ERR_CUDA;
    }
    unsigned long long mxb = 0;
    long long mnb = 0;
    const int rc2 = st_lane_read_status(lane, lane->streams[0], &mxb, &mnb);  // also clears the word for the next call
    if (rc != ST_OK) return rc;
    if (rc2 != ST_OK) return rc2;
    if (mxb != 0 || mnb != 0) {
        // the reference reports max_id when it is >= size, else min_id (MuchTree.pyx:1303-1310)
        st_set_bad_node(mxb != 0 ? (int64_t)mxb : (int64_t)mnb);
        st_set_error("node id %lld out of bounds (tree size %lld)", (long long)st_bad_node(), (long long)t->n_nodes);
        return ST_ERR_NODE_RANGE;
    }
    return ST_OK;
}
