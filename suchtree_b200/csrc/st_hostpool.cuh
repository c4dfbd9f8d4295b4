// Small persistent host thread pool for the drop-in (host-buffer) entry points:
// packs the caller's int64 id arrays into int32 pinned staging (half the PCIe
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
