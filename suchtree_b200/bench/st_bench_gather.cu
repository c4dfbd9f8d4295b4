// Bench-only library (libsuchtree_b200_bench.so): the L2 random-sector gather
// micro-benchmark that supplies the gather roofline of SURVEY.md §8d, and the experiment
// kernels that compare the hardware paths a random sector can take (LSU, texture, bulk
// copy, tensor-map gather4, mixtures; SUCHTREE_B200_GATHER_MODE).  Measurement tooling, not
// product: nothing in libsuchtree_b200.so depends on it.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../csrc/st_device.cuh"
#include "suchtree_b200_bench.h"

// this library's own error text (the product library's st_set_error is not linked here)
static thread_local char g_bench_err[512] = "";
void st_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_bench_err, sizeof(g_bench_err), fmt, ap);
    va_end(ap);
}
void st_set_bad_node(int64_t) {}
extern "C" const char *st_bench_last_error(void) { return g_bench_err; }

// ------------------------------------------------------ gather roofline -----
// Every thread issues `loads` independent 32-byte (one sector) loads at
// pseudo-random sector addresses of a `bytes`-sized buffer, 4 in flight at a
// time -- the access pattern of the pair kernel's index lookups.
__global__ void __launch_bounds__(512)
k_gather(const ulonglong4 *__restrict__ buf, uint64_t n_sectors, int64_t loads, uint64_t seed,
         unsigned long long *__restrict__ sink) {
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (int64_t i = 0; i < loads; i += 4) {
        Philox4 r = st_philox4x32_10(tid * uint64_t(loads) + uint64_t(i), seed);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        uint64_t a[4], b[4], c[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
            asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(a[k]), "=l"(b[k]), "=l"(c[k]), "=l"(d[k])
                         : "l"(buf + idx));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc ^= a[k] ^ b[k] ^ c[k] ^ d[k];
    }
    if (acc == 0x123456789abcdefull) *sink = acc;  // keep the loads alive
}

// Experiment: the same random sectors fetched by the bulk-copy (TMA) engine
// straight into shared memory (cp.async.bulk, one 32-byte copy per sector,
// completion on an mbarrier), bypassing the LSU/L1TEX data pipe.
// Selected with SUCHTREE_B200_GATHER_MODE=bulk.
__global__ void __launch_bounds__(256)
k_gather_bulk(const ulonglong4 *__restrict__ buf, uint64_t n_sectors, int64_t loads, uint64_t seed,
              unsigned long long *__restrict__ sink) {
    __shared__ __align__(128) ulonglong4 slots[256 * 4];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    uint32_t phase = 0;
    for (int64_t i = 0; i < loads; i += 4) {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                         "r"(256 * 4 * 32)
                         : "memory");
        Philox4 r = st_philox4x32_10(tid * uint64_t(loads) + uint64_t(i), seed);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slots[threadIdx.x * 4 + k]);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(dst),
                "l"(buf + idx), "r"(bar_a)
                : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar_a), "r"(phase)
                : "memory");
        }
        phase ^= 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ulonglong4 v = slots[threadIdx.x * 4 + k];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncthreads();  // slots are rewritten next round
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// Experiment: route the sector gathers through the TEXTURE data pipe
// (tex1Dfetch on a linear int4 texture, two 16-byte texels per 32-byte sector).
// MODE 0: all four gathers of a round via TEX; MODE 1: two via TEX, two via LSU.
// Selected with SUCHTREE_B200_GATHER_MODE=tex / mix.
template <int MODE>
__global__ void __launch_bounds__(512)
k_gather_tex(cudaTextureObject_t tex, const ulonglong4 *__restrict__ buf, uint64_t n_sectors,
             int64_t loads, uint64_t seed, unsigned long long *__restrict__ sink) {
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (int64_t i = 0; i < loads; i += 4) {
        Philox4 r = st_philox4x32_10(tid * uint64_t(loads) + uint64_t(i), seed);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        int4 ta[4], tb[4];
        uint64_t a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
            if (MODE == 0 || k < 2) {
                ta[k] = tex1Dfetch<int4>(tex, int(2 * idx));
                tb[k] = tex1Dfetch<int4>(tex, int(2 * idx + 1));
            } else {
                ta[k] = tb[k] = make_int4(0, 0, 0, 0);
                asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                             : "=l"(a[k]), "=l"(b[k]), "=l"(c[k]), "=l"(d[k])
                             : "l"(buf + idx));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc ^= a[k] ^ b[k] ^ c[k] ^ d[k] ^ uint64_t(uint32_t(ta[k].x ^ ta[k].y ^ ta[k].z ^ ta[k].w)) ^
                   (uint64_t(uint32_t(tb[k].x ^ tb[k].y ^ tb[k].z ^ tb[k].w)) << 32);
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}


// Experiment: TMA tile::gather4 (sm_100): ONE bulk-tensor instruction fetches four
// rows (here: four 32-byte sectors) of a 2-D view of the buffer into shared memory.
// MIX = 0: every sector via gather4.  MIX = K: per round each thread also issues
// 4*K LSU sector loads, to see whether the TMA path adds to the L1TEX gather rate.
// Selected with SUCHTREE_B200_GATHER_MODE=g4 / g4mix1 / g4mix2.
#include <cuda.h>

template <int MIX>
__global__ void __launch_bounds__(128)
k_gather_g4(const __grid_constant__ CUtensorMap tmap, const ulonglong4 *__restrict__ buf,
            uint64_t n_sectors, int64_t rounds, uint64_t seed, unsigned long long *__restrict__ sink) {
    constexpr int STAGES = 2;
    __shared__ __align__(128) ulonglong4 slots[STAGES][128 * 4];
    __shared__ __align__(8) uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    auto issue = [&](int64_t r) {
        const int s = int(r % STAGES);
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[s]);
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(128 * 4 * 32) : "memory");
        Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed);
        int32_t r0 = int32_t((uint64_t(p.x) * n_sectors) >> 32), r1 = int32_t((uint64_t(p.y) * n_sectors) >> 32);
        int32_t r2 = int32_t((uint64_t(p.z) * n_sectors) >> 32), r3 = int32_t((uint64_t(p.w) * n_sectors) >> 32);
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slots[s][threadIdx.x * 4]);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(dst), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
            : "memory");
    };
    issue(0);
    for (int64_t r = 0; r < rounds; ++r) {
        if (r + 1 < rounds) issue(r + 1);
        if (MIX > 0) {
#pragma unroll
            for (int m = 0; m < MIX; ++m) {
                Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed + 77 + m);
                const uint32_t w[4] = {p.x, p.y, p.z, p.w};
                uint64_t a[4], b[4], c[4], d[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
                    asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                                 : "=l"(a[k]), "=l"(b[k]), "=l"(c[k]), "=l"(d[k])
                                 : "l"(buf + idx));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) acc ^= a[k] ^ b[k] ^ c[k] ^ d[k];
            }
        }
        const int s = int(r % STAGES);
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[s]);
        const uint32_t phase = uint32_t(r / STAGES) & 1;
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ulonglong4 v = slots[s][threadIdx.x * 4 + k];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncthreads();  // stage s is rewritten by issue(r + 2)
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// Experiment: warp-specialised mix.  Warps [0, TMA_WARPS) of every 16-warp CTA fetch
// sectors with gather4 (per-warp mbarrier, two stages, no CTA-wide sync); the other
// warps run the plain LSU gather loop.  Are the two paths into L2 additive?
// Selected with SUCHTREE_B200_GATHER_MODE=ws<k> (k = TMA warps per CTA, 0..16).
template <int TMA_WARPS>
__global__ void __launch_bounds__(512)
k_gather_ws(const __grid_constant__ CUtensorMap tmap, const ulonglong4 *__restrict__ buf,
            uint64_t n_sectors, int64_t rounds, uint64_t seed, unsigned long long *__restrict__ sink) {
    constexpr int TW = TMA_WARPS > 0 ? TMA_WARPS : 1;
    __shared__ __align__(128) ulonglong4 slots[TW][2][32 * 4];  // 8 KB per TMA warp
    __shared__ __align__(8) uint64_t bar[TW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    if (warp < TMA_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[warp][s])), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        auto issue = [&](int64_t r) {
            const int s = int(r & 1);
            const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[warp][s]);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32 * 4 * 32) : "memory");
            __syncwarp();
            Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed);
            int32_t r0 = int32_t((uint64_t(p.x) * n_sectors) >> 32), r1 = int32_t((uint64_t(p.y) * n_sectors) >> 32);
            int32_t r2 = int32_t((uint64_t(p.z) * n_sectors) >> 32), r3 = int32_t((uint64_t(p.w) * n_sectors) >> 32);
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slots[warp][s][lane * 4]);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                ::"r"(dst), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
                : "memory");
        };
        issue(0);
        for (int64_t r = 0; r < rounds; ++r) {
            if (r + 1 < rounds) issue(r + 1);
            const int s = int(r & 1);
            const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[warp][s]);
            const uint32_t phase = uint32_t(r >> 1) & 1;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ulonglong4 v = slots[warp][s][lane * 4 + k];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();  // stage s is rewritten by issue(r + 2)
        }
    } else {
        for (int64_t r = 0; r < rounds; ++r) {
            Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed);
            const uint32_t w[4] = {p.x, p.y, p.z, p.w};
            uint64_t a[4], b[4], c[4], d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
                asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                             : "=l"(a[k]), "=l"(b[k]), "=l"(c[k]), "=l"(d[k])
                             : "l"(buf + idx));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) acc ^= a[k] ^ b[k] ^ c[k] ^ d[k];
        }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// Experiment: are the LSU and the TMA gather paths additive?  Every CTA has 16 LSU warps
// (the plain gather loop, `rounds` rounds of 4 sectors) PLUS `TW` TMA warps that fetch
// sectors with gather4 for `rounds_tma` rounds (per-warp mbarriers, two stages, dynamic
// shared memory).  The host reports (LSU sectors + TMA sectors) / kernel time; TW = 0 with
// the same CTA shape is the baseline.  SUCHTREE_B200_GATHER_MODE=add<TW>_<percent>.
__global__ void __launch_bounds__(1024)
k_gather_add(const __grid_constant__ CUtensorMap tmap, const ulonglong4 *__restrict__ buf,
             uint64_t n_sectors, int64_t rounds, int64_t rounds_tma, int tma_warps, uint64_t seed,
             unsigned long long *__restrict__ sink) {
    extern __shared__ __align__(128) unsigned char dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    if (warp >= 16) {
        const int w = warp - 16;
        ulonglong4 *slots = reinterpret_cast<ulonglong4 *>(dyn) + size_t(w) * 2 * 128;  // [2][128]
        uint64_t *bar = reinterpret_cast<uint64_t *>(dyn + size_t(tma_warps) * 8192) + 2 * w;
        if (lane == 0) {
            for (int s = 0; s < 2; ++s)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[s])), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        auto issue = [&](int64_t r) {
            const int s = int(r & 1);
            const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[s]);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32 * 4 * 32) : "memory");
            __syncwarp();
            Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed);
            int32_t r0 = int32_t((uint64_t(p.x) * n_sectors) >> 32), r1 = int32_t((uint64_t(p.y) * n_sectors) >> 32);
            int32_t r2 = int32_t((uint64_t(p.z) * n_sectors) >> 32), r3 = int32_t((uint64_t(p.w) * n_sectors) >> 32);
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slots[s * 128 + lane * 4]);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                ::"r"(dst), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
                : "memory");
        };
        if (rounds_tma > 0) issue(0);
        for (int64_t r = 0; r < rounds_tma; ++r) {
            if (r + 1 < rounds_tma) issue(r + 1);
            const int s = int(r & 1);
            const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[s]);
            const uint32_t phase = uint32_t(r >> 1) & 1;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            }
            // one 32-byte read per lane and round: consumers of a real kernel would read their
            // records from shared memory like this
            ulonglong4 v = slots[s * 128 + lane * 4];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
            __syncwarp();
        }
    } else {
        for (int64_t r = 0; r < rounds; ++r) {
            Philox4 p = st_philox4x32_10(tid * uint64_t(rounds) + uint64_t(r), seed);
            const uint32_t w[4] = {p.x, p.y, p.z, p.w};
            uint64_t a[4], b[4], c[4], d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t idx = (uint64_t(w[k]) * n_sectors) >> 32;
                asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                             : "=l"(a[k]), "=l"(b[k]), "=l"(c[k]), "=l"(d[k])
                             : "l"(buf + idx));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) acc ^= a[k] ^ b[k] ^ c[k] ^ d[k];
        }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// 2-D view of a buffer of 32-byte sectors: inner dim = 8 x u32, outer dim = sectors
static int make_sector_tmap(void *buf, uint64_t n_sectors, CUtensorMap *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        st_set_error("cuTensorMapEncodeTiled not available");
        return ST_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {8, n_sectors};
    cuuint64_t gstride[1] = {32};
    cuuint32_t box[2] = {8, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, buf, gdim, gstride, box,
                                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        st_set_error("cuTensorMapEncodeTiled failed: %d", int(r));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}

extern "C" int st_bench_gather(int device, int64_t bytes, int64_t loads_per_thread, int iters,
                               double *sectors_per_s) {
    if (!sectors_per_s || bytes < 32 || loads_per_thread < 4 || iters < 1) {
        st_set_error("st_bench_gather: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        st_set_error("st_bench_gather: no CUDA device %d", device);
        return ST_ERR_CUDA;
    }
    DeviceGuard g(device);
    cudaDeviceProp prop;
    ST_CUDA(cudaGetDeviceProperties(&prop, device));
    const uint64_t n_sectors = uint64_t(bytes) / 32;
    void *buf = nullptr;
    unsigned long long *sink = nullptr;
    ST_CUDA(cudaMalloc(&buf, n_sectors * 32));
    ST_CUDA(cudaMalloc(&sink, 8));
    ST_CUDA(cudaMemset(buf, 1, n_sectors * 32));
    const int grid = prop.multiProcessorCount * 4, tpb = 512;
    loads_per_thread = (loads_per_thread + 3) / 4 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *mode = getenv("SUCHTREE_B200_GATHER_MODE");
    const bool bulk = mode && mode[0] == 'b';
    const bool use_tex = mode && mode[0] == 't', mix = mode && mode[0] == 'm';
    const bool g4 = mode && (mode[0] == 'g' || mode[0] == 'w' || mode[0] == 'a');
    int add_tw = -1, add_pct = 0;
    if (mode && mode[0] == 'a') sscanf(mode, "add%d_%d", &add_tw, &add_pct);
    const int ws = (mode && mode[0] == 'w') ? atoi(mode + 2) : -1;
    const int g4mix = (g4 && mode[2] == 'm') ? (mode[5] ? mode[5] - '0' : 1) : 0;
    CUtensorMap tmap;
    if (g4) {
        int rc = make_sector_tmap(buf, n_sectors, &tmap);
        if (rc != ST_OK) return rc;
    }
    cudaTextureObject_t tex = 0;
    if (use_tex || mix) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = buf;
        rd.res.linear.desc = cudaCreateChannelDesc<int4>();
        rd.res.linear.sizeInBytes = n_sectors * 32;
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        ST_CUDA(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    }
    auto launch = [&](uint64_t seed) {
        if (use_tex)
            k_gather_tex<0><<<grid, tpb>>>(tex, static_cast<const ulonglong4 *>(buf), n_sectors, loads_per_thread, seed, sink);
        else if (mix)
            k_gather_tex<1><<<grid, tpb>>>(tex, static_cast<const ulonglong4 *>(buf), n_sectors, loads_per_thread, seed, sink);
        else if (add_tw >= 0) {
            const int64_t rounds = loads_per_thread / 4, rounds_tma = rounds * add_pct / 100;
            const int threads = 512 + 32 * add_tw, smem = add_tw * 8192 + add_tw * 16 + 16;
            cudaFuncSetAttribute(k_gather_add, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            k_gather_add<<<prop.multiProcessorCount * 3, threads, smem>>>(tmap, static_cast<const ulonglong4 *>(buf), n_sectors,
                                                                       rounds, rounds_tma, add_tw, seed, sink);
        } else if (ws >= 0) {
            const int64_t rounds = loads_per_thread / 4;
            const ulonglong4 *b = static_cast<const ulonglong4 *>(buf);
            switch (ws) {  // 3 CTAs of 512 threads per SM at most (shared memory of the TMA warps)
                case 0: k_gather_ws<0><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
                case 2: k_gather_ws<2><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
                case 4: k_gather_ws<4><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
                case 1: k_gather_ws<1><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
                case 3: k_gather_ws<3><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
                default: k_gather_ws<5><<<grid, tpb>>>(tmap, b, n_sectors, rounds, seed, sink); break;
            }
        } else if (g4) {
            const int64_t rounds = loads_per_thread / 4;
            if (g4mix == 0) k_gather_g4<0><<<grid * 4, tpb / 4>>>(tmap, static_cast<const ulonglong4 *>(buf), n_sectors, rounds, seed, sink);
            else if (g4mix == 1) k_gather_g4<1><<<grid * 4, tpb / 4>>>(tmap, static_cast<const ulonglong4 *>(buf), n_sectors, rounds, seed, sink);
            else if (g4mix == 2) k_gather_g4<2><<<grid * 4, tpb / 4>>>(tmap, static_cast<const ulonglong4 *>(buf), n_sectors, rounds, seed, sink);
            else k_gather_g4<3><<<grid * 4, tpb / 4>>>(tmap, static_cast<const ulonglong4 *>(buf), n_sectors, rounds, seed, sink);
        } else if (bulk)  // same thread count: twice the CTAs of half the size
            k_gather_bulk<<<grid * 2, tpb / 2>>>(static_cast<const ulonglong4 *>(buf), n_sectors, loads_per_thread, seed, sink);
        else
            k_gather<<<grid, tpb>>>(static_cast<const ulonglong4 *>(buf), n_sectors, loads_per_thread, seed, sink);
    };
    launch(1);  // warm
    float best_ms = 1e30f;
    for (int it = 0; it < iters; ++it) {
        cudaEventRecord(e0);
        launch(uint64_t(it) + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best_ms = std::min(best_ms, ms);
    }
    cudaError_t e = cudaGetLastError();
    if (tex) cudaDestroyTextureObject(tex);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    if (e != cudaSuccess) {
        st_set_error("st_bench_gather: %s", cudaGetErrorString(e));
        return ST_ERR_CUDA;
    }
    if (add_tw >= 0) {
        const double rounds = double(loads_per_thread / 4), rounds_tma = double((loads_per_thread / 4) * add_pct / 100);
        const double per_cta = 512.0 * rounds * 4.0 + 32.0 * add_tw * rounds_tma * 4.0;
        *sectors_per_s = double(prop.multiProcessorCount) * 3.0 * per_cta / (double(best_ms) * 1e-3);
        return ST_OK;
    }
    *sectors_per_s = double(grid) * tpb * double(loads_per_thread) * (1 + (ws >= 0 ? 0 : g4mix)) / (double(best_ms) * 1e-3);
    return ST_OK;
}
