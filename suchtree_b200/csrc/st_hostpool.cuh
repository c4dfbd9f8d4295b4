// Small persistent host thread pool for the drop-in (host-buffer) entry points:
// packs the caller's int64 id arrays into pinned staging (a bit stream of 2 x id_bits per pair
// for the chunked pair pipeline, int32 for medium-size calls and quartets: a fraction of the PCIe
// bytes) and copies results out of pinned staging, in parallel, while the GPU
// works on the previous chunk.
#pragma once

#include <cstdint>
#include <functional>

// runs fn(part, n_parts) for part = 0..n_parts-1 on the pool (the caller takes
// part 0) and returns when all parts are done.  Safe to call from several
// threads (calls serialise).  SUCHTREE_B200_HOST_THREADS caps the pool size.
void st_parallel_for(int n_parts, const std::function<void(int, int)> &fn);
int st_host_threads();  // parts worth using for bandwidth-bound loops

// int64 ids -> int32 into (pinned) staging, in parallel, streaming stores; `width` ids per row
// (2: pairs, 4: quartets), rows strided by s0, columns by s1 (elements).  Returns the OR of
// everything seen: any bit >= 31 set <=> some id is negative or >= 2^31.
uint64_t st_pack_ids(const int64_t *src, int64_t s0, int64_t s1, int64_t rows, int32_t *dst, int width);
// int64 pairs -> the bit-packed pair stream of the pair kernel (2w bits per pair, a low, b high);
// returns the OR of all ids seen; dst: ceil(rows * 2w / 64) + 1 words
uint64_t st_pack_pairs_bits(const int64_t *src, int64_t s0, int64_t s1, int64_t rows, uint64_t *dst, int w);
void st_parallel_copy(void *dst, const void *src, size_t bytes);
// the reference's report for an out-of-range array (MuchTree.pyx:897-903): max id if it is
// >= n_nodes, else min id -> st_bad_node() / st_last_error()
void st_report_range(const int64_t *src, int64_t s0, int64_t s1, int64_t rows, int width, int64_t n_nodes);
