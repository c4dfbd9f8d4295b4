// Experiment: what does it cost to get a page-locked block of `bytes` bytes?
//   a  cudaHostAlloc
//   b  mmap + MADV_HUGEPAGE + parallel first touch + cudaHostRegister
//   c  mmap (4 KiB pages) + parallel first touch + cudaHostRegister
//   d  as b, registering in 64 MiB slices (a pipeline could start copying after the first)
// Build: nvcc -O2 -o scripts/_bin/pin_exp scripts/pin_exp.cu ; run under gpurun.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void touch(char *p, size_t bytes, int threads) {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=] {
            const size_t b = bytes * t / threads, e = bytes * (t + 1) / threads;
            for (size_t o = b; o < e; o += 4096) p[o] = 1;
        });
    for (auto &x : th) x.join();
}
static double d2h_rate(void *h, size_t bytes) {
    void *d = nullptr;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    cudaDeviceSynchronize();
    const double t0 = now();
    cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
    const double dt = now() - t0;
    cudaFree(d);
    return bytes / dt / 1e9;
}

int main(int argc, char **argv) {
    const size_t bytes = argc > 1 ? size_t(atoll(argv[1])) : size_t(800) << 20;
    const int threads = argc > 2 ? atoi(argv[2]) : 16;
    cudaFree(0);
    {
        void *w = nullptr;  // warm the driver's pinning path
        cudaHostAlloc(&w, 1 << 20, cudaHostAllocPortable);
        cudaFreeHost(w);
    }
    for (int rep = 0; rep < 2; ++rep) {
        {
            void *p = nullptr;
            const double t0 = now();
            cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
            const double dt = now() - t0;
            printf("a cudaHostAlloc            %.3f s  (%s)  d2h %.1f GB/s\n", dt, cudaGetErrorString(e), d2h_rate(p, bytes));
            cudaFreeHost(p);
        }
        for (int huge = 1; huge >= 0; --huge) {
            const double t0 = now();
            char *p = static_cast<char *>(mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
            char *q = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + (2 << 20) - 1) & ~uintptr_t((2 << 20) - 1));
            int mr = huge ? madvise(q, bytes, MADV_HUGEPAGE) : madvise(q, bytes, MADV_NOHUGEPAGE);
            const double t1 = now();
            touch(q, bytes, threads);
            const double t2 = now();
            cudaError_t e = cudaHostRegister(q, bytes, cudaHostRegisterPortable);
            const double t3 = now();
            printf("%c mmap huge=%d madvise=%d   map %.3f touch %.3f register %.3f total %.3f s (%s)  d2h %.1f GB/s\n",
                   huge ? 'b' : 'c', huge, mr, t1 - t0, t2 - t1, t3 - t2, t3 - t0, cudaGetErrorString(e), d2h_rate(q, bytes));
            cudaHostUnregister(q);
            munmap(p, bytes + (2 << 20));
        }
        {
            const double t0 = now();
            char *p = static_cast<char *>(mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
            char *q = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + (2 << 20) - 1) & ~uintptr_t((2 << 20) - 1));
            madvise(q, bytes, MADV_HUGEPAGE);
            touch(q, bytes, threads);
            const double t1 = now();
            const size_t slice = size_t(64) << 20;
            double first = 0;
            for (size_t o = 0; o < bytes; o += slice) {
                cudaHostRegister(q + o, bytes - o < slice ? bytes - o : slice, cudaHostRegisterPortable);
                if (o == 0) first = now() - t1;
            }
            const double t2 = now();
            printf("d huge, 64 MiB slices      touch %.3f first slice %.3f all slices %.3f total %.3f s\n", t1 - t0, first,
                   t2 - t1, t2 - t0);
            for (size_t o = 0; o < bytes; o += slice) cudaHostUnregister(q + o);
            munmap(p, bytes + (2 << 20));
        }
    }
    // a numpy-like allocation: malloc (glibc mmap above the threshold) + MADV_HUGEPAGE as numpy does
    {
        const double t0 = now();
        char *p = static_cast<char *>(malloc(bytes));
        madvise(reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(p) + 4095) & ~uintptr_t(4095)), bytes - 4096, MADV_HUGEPAGE);
        touch(p, bytes, 1);
        const double t1 = now();
        cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
        const double t2 = now();
        printf("e malloc+madvise, 1 thread touch %.3f register %.3f (%s)\n", t1 - t0, t2 - t1, cudaGetErrorString(e));
        cudaHostUnregister(p);
        free(p);
    }
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) {
        char buf[128] = {0};
        if (fgets(buf, sizeof buf, f)) printf("THP: %s", buf);
        fclose(f);
    }
    return 0;
}
