// pearson() on host vectors.
//
// st_pearson replaces _pearson (MuchTree.pyx:62-79): the same two-pass formula
// (means first, then centred sums, +1e-20 guard) with fp64 accumulators and a
// fixed reduction tree instead of the reference's sequential fp32 accumulators.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "st_device.cuh"
#include "st_hostctx.cuh"
#include "st_hostpool.cuh"

static const int PT = 256;

__device__ __forceinline__ double wsum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int K>
__device__ __forceinline__ void block_partials(double (&v)[K], double *__restrict__ partials) {
    __shared__ double red[K][PT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = wsum(v[k]);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0;
        for (int w = 0; w < PT / 32; ++w) s += red[threadIdx.x][w];
        partials[size_t(blockIdx.x) * K + threadIdx.x] = s;
    }
}

// five shifted moments of one chunk: sum dx, sum dy, sum dx^2, sum dy^2, sum dx dy with
// dx = x - x0, dy = y - y0 (x0, y0 close to the means: no cancellation in the centred sums)
__global__ void __launch_bounds__(PT)
k_pearson_moments(int64_t n, const double *__restrict__ x, const double *__restrict__ y, double x0, double y0,
                  double *__restrict__ partials) {
    double v[5] = {0, 0, 0, 0, 0};
    for (int64_t i = int64_t(blockIdx.x) * PT + threadIdx.x; i < n; i += int64_t(gridDim.x) * PT) {
        const double xt = x[i] - x0, yt = y[i] - y0;
        v[0] += xt;
        v[1] += yt;
        v[2] += xt * xt;
        v[3] += yt * yt;
        v[4] += xt * yt;
    }
    block_partials<5>(v, partials);
}

template <int K>
__global__ void k_fold(int grid, const double *__restrict__ partials, double *__restrict__ out) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= K) return;
    double s = 0;
    for (int i = lane; i < grid; i += 32) s += partials[size_t(i) * K + k];
    s = wsum(s);
    if (lane == 0) out[k] = s;
}

// st_pearson streams the two vectors through the lane: chunk c of x and y is brought to the device
// (straight DMA when the caller's memory is page-locked, as the result arrays of linked_distances()
// are; else the host pool copies it into the lane's page-locked staging first, under the DMA of
// chunk c-1) and folded into five shifted moments; nothing of size n is allocated anywhere.  The
// shift is the mean of the first few thousand elements, so the one pass loses nothing against the
// reference's two passes (MuchTree.pyx:62-79 first subtracts the exact means); partials are
// folded in a fixed order.
extern "C" int st_pearson(int device, const double *x, const double *y, int64_t n, double *r) {
    if (!r || n < 0 || (n > 0 && (!x || !y))) {
        st_set_error("st_pearson: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    if (n == 0) {  // the reference's ZeroDivisionError branch returns 0.0 (MuchTree.pyx:86-87)
        *r = 0.0;
        return ST_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        st_set_error("st_pearson: no CUDA device %d (no CPU fallback)", device);
        return ST_ERR_CUDA;
    }
    DeviceGuard g(device);
    LaneGuard lg(device);
    HostLane *lane = lg.lane;
    if (!lane) return ST_ERR_CUDA;
    int sms = 0;
    ST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const bool pinned = st_is_pinned(x) && st_is_pinned(x + n - 1) && st_is_pinned(y) && st_is_pinned(y + n - 1);
    int rc = st_lane_ensure_stage(lane, std::min<int64_t>(n, ST_STAGE_PAIRS_MAX), !pinned, false);
    if (rc != ST_OK) return rc;
    const int64_t C = lane->stage_pairs;  // elements of x and of y per chunk: 8 + 8 of the slot's 16 bytes per pair
    const int64_t n_chunks = (n + C - 1) / C;
    const int grid = int(std::min<int64_t>((std::min(n, C) + PT - 1) / PT, int64_t(sms) * 8));
    double x0 = 0.0, y0 = 0.0;
    const int64_t pilot = std::min<int64_t>(n, 4096);
    for (int64_t i = 0; i < pilot; ++i) {
        x0 += x[i];
        y0 += y[i];
    }
    x0 /= double(pilot);
    y0 /= double(pilot);
    if (!std::isfinite(x0)) x0 = 0.0;
    if (!std::isfinite(y0)) y0 = 0.0;
    cudaStream_t s = lane->streams[0];
    double *dp = nullptr;
    if (cudaMallocAsync(reinterpret_cast<void **>(&dp), size_t(n_chunks) * grid * 5 * 8, s) != cudaSuccess) {
        cudaGetLastError();
        st_set_error("st_pearson: device allocation failed");
        return ST_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        cudaFreeAsync(dp, s);
        cudaStreamSynchronize(s);
        return code;
    };
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int k = int(c % ST_LANE_SLOTS);
        const int64_t off = c * C, m = std::min(C, n - off);
        double *dx = static_cast<double *>(lane->d_in[k]), *dy = dx + C;
        // the slot's previous chunk (this call's or an earlier one's) has left its buffers
        if (cudaEventSynchronize(lane->ev[k]) != cudaSuccess) return fail(ST_ERR_CUDA);
        const double *hx = x + off, *hy = y + off;
        if (!pinned) {
            double *sx = static_cast<double *>(lane->h_in[k]), *sy = sx + C;
            st_parallel_copy(sx, hx, size_t(m) * 8);
            st_parallel_copy(sy, hy, size_t(m) * 8);
            hx = sx;
            hy = sy;
        }
        if (cudaMemcpyAsync(dx, hx, size_t(m) * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(dy, hy, size_t(m) * 8, cudaMemcpyHostToDevice, s) != cudaSuccess)
            return fail(ST_ERR_CUDA);
        k_pearson_moments<<<grid, PT, 0, s>>>(m, dx, dy, x0, y0, dp + size_t(c) * grid * 5);
        if (cudaEventRecord(lane->ev[k], s) != cudaSuccess) return fail(ST_ERR_CUDA);
    }
    double *ds = lane->d_scratch, *h = lane->h_scratch;
    k_fold<5><<<1, 160, 0, s>>>(int(n_chunks) * grid, dp, ds);
    cudaMemcpyAsync(h, ds, 5 * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaFreeAsync(dp, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        st_set_error("st_pearson: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    // centred sums from the shifted ones (the shift cancels exactly in exact arithmetic)
    const double nn = double(n);
    const double cxx = h[2] - h[0] * h[0] / nn, cyy = h[3] - h[1] * h[1] / nn, cxy = h[4] - h[0] * h[1] / nn;
    *r = cxy / std::sqrt(cxx * cyy + 1.0e-20);
    return ST_OK;
}
