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

__global__ void __launch_bounds__(PT)
k_pearson_means(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                double *__restrict__ partials) {
    double v[2] = {0, 0};
    for (int64_t i = int64_t(blockIdx.x) * PT + threadIdx.x; i < n; i += int64_t(gridDim.x) * PT) {
        v[0] += x[i];
        v[1] += y[i];
    }
    block_partials<2>(v, partials);
}

__global__ void __launch_bounds__(PT)
k_pearson_centred(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                  const double *__restrict__ sums, double *__restrict__ partials) {
    const double ax = sums[0] / double(n), ay = sums[1] / double(n);
    double v[3] = {0, 0, 0};
    for (int64_t i = int64_t(blockIdx.x) * PT + threadIdx.x; i < n; i += int64_t(gridDim.x) * PT) {
        double xt = x[i] - ax, yt = y[i] - ay;
        v[0] += xt * xt;
        v[1] += yt * yt;
        v[2] += xt * yt;
    }
    block_partials<3>(v, partials);
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
    cudaDeviceProp prop;
    ST_CUDA(cudaGetDeviceProperties(&prop, device));
    int grid = int(std::min<int64_t>((n + PT - 1) / PT, int64_t(prop.multiProcessorCount) * 8));
    double *dx = nullptr, *dy = nullptr, *dp = nullptr, *ds = nullptr;
    auto cleanup = [&]() {
        cudaFree(dx);
        cudaFree(dy);
        cudaFree(dp);
        cudaFree(ds);
    };
    if (cudaMalloc(&dx, size_t(n) * 8) != cudaSuccess || cudaMalloc(&dy, size_t(n) * 8) != cudaSuccess ||
        cudaMalloc(&dp, size_t(grid) * 3 * 8) != cudaSuccess || cudaMalloc(&ds, 5 * 8) != cudaSuccess) {
        cleanup();
        st_set_error("st_pearson: cudaMalloc failed");
        return ST_ERR_NOMEM;
    }
    cudaStream_t s = nullptr;
    cudaMemcpyAsync(dx, x, size_t(n) * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(dy, y, size_t(n) * 8, cudaMemcpyHostToDevice, s);
    k_pearson_means<<<grid, PT, 0, s>>>(n, dx, dy, dp);
    k_fold<2><<<1, 64, 0, s>>>(grid, dp, ds);
    k_pearson_centred<<<grid, PT, 0, s>>>(n, dx, dy, ds, dp);
    k_fold<3><<<1, 96, 0, s>>>(grid, dp, ds + 2);
    double h[5];
    cudaMemcpyAsync(h, ds, sizeof(h), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) {
        st_set_error("st_pearson: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    *r = h[4] / std::sqrt(h[2] * h[3] + 1.0e-20);
    return ST_OK;
}
