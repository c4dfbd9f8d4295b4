/* suchtree_b200 -- bench-only library (libsuchtree_b200_bench.so).
 *
 * Measurement tooling that is NOT part of the drop-in boundary: the L2 random-sector
 * gather micro-benchmark behind bench.py's `gather_roofline` (SURVEY.md section 8d asks for
 * it: no gather figure exists in MEASURED_PEAKS.json) and the experiment kernels comparing
 * the hardware paths a random sector can take (SUCHTREE_B200_GATHER_MODE = bulk | tex | mix |
 * g4... | ws... | add..., see suchtree_b200/bench/st_bench_gather.cu). */
#ifndef SUCHTREE_B200_BENCH_H
#define SUCHTREE_B200_BENCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ST_BENCH_API __attribute__((visibility("default")))
#else
#define ST_BENCH_API
#endif

/* random 32-byte-sector gather rate over a `bytes`-sized device buffer: every thread
 * issues loads_per_thread independent one-sector loads at Philox-random sector addresses,
 * four in flight at a time (the access pattern of the pair kernel's record lookups);
 * best of `iters` launches, CUDA events. */
ST_BENCH_API int st_bench_gather(int device, int64_t bytes, int64_t loads_per_thread, int iters,
                                 double *sectors_per_s);
ST_BENCH_API const char *st_bench_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SUCHTREE_B200_BENCH_H */
