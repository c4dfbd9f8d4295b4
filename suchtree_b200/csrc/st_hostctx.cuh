// Per-device host context shared by every host-buffer entry point (st_distances,
// st_mrca, st_quartet_topologies, st_distance_matrix with host output, the linked /
// sampler calls): a small pool of LANES.  A lane is everything one in-flight host call
// needs -- three streams + events, device and pinned staging, and its OWN range-status
// word -- so concurrent host callers run side by side instead of queueing on a per-tree
// mutex, and no caller can see (or clear) another caller's out-of-range flag.
// Staging is per device, not per tree: it is allocated once (in the background, when the
// first tree is created on the device), not at the first large query of every tree.
#pragma once

#include "st_internal.cuh"

static const int ST_LANE_SLOTS = 3;
static const int64_t ST_STAGE_PAIRS_MAX = int64_t(1) << 22;  // pairs per chunk of the host pipeline
static const int64_t ST_SMALL_CALL = 4096;      // host-side range check, scalar pack
static const int64_t ST_MEDIUM_CALL = 262144;   // pool pack, device-side range check, still one zero-copy kernel

struct HostLane {
    int device = 0;
    cudaStream_t streams[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};
    void *d_in[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};    // 16 B per pair
    void *d_out[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};   //  8 B per pair
    void *d_out2[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};  //  4 B per pair
    void *h_in[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};    // pinned, 16 B per pair
    void *h_out[ST_LANE_SLOTS] = {nullptr, nullptr, nullptr};   // pinned,  8 B per pair
    int64_t stage_pairs = 0;     // capacity of the staging above, in pairs
    RangeStatus *d_status = nullptr;  // this lane's status word (device), zero between calls
    RangeStatus *h_status = nullptr;  // pinned landing zone for it
    void *h_small_in = nullptr, *h_small_out = nullptr;  // pinned, ST_MEDIUM_CALL pairs: the zero-copy latency path
    double *d_scratch = nullptr;      // 64 doubles: moment sums (the all-reduce runs in place here)
    double *h_scratch = nullptr;      // pinned landing zone for them
    bool busy = false;
};

// Blocks until a lane of `device` is free (creating one if fewer than the cap exist).
// NULL + st_last_error() on CUDA failure.  The caller must have made `device` current.
HostLane *st_lane_acquire(int device);
void st_lane_release(HostLane *lane);

struct LaneGuard {
    HostLane *lane;
    explicit LaneGuard(int device) : lane(st_lane_acquire(device)) {}
    ~LaneGuard() {
        if (lane) st_lane_release(lane);
    }
    LaneGuard(const LaneGuard &) = delete;
    LaneGuard &operator=(const LaneGuard &) = delete;
};

// grow the lane's staging to hold chunks of min(n, ST_STAGE_PAIRS_MAX) pairs
int st_lane_ensure_stage(HostLane *lane, int64_t n, bool need_h_in, bool need_h_out);

// page-locked memory from the pool behind st_host_alloc (release with st_host_free)
int st_pinned_alloc(size_t bytes, void **out);

// Read the lane's status word after its kernels (synchronises `stream`, which must be
// ordered after every kernel that could have written it); clears it when set.
// *max_bad / *min_bad = 0 when nothing was flagged.
int st_lane_read_status(HostLane *lane, cudaStream_t stream, unsigned long long *max_bad, long long *min_bad);

// start allocating lane 0's staging for `device` on a helper thread (once per device)
void st_hostctx_prewarm(int device);

// cudaPointerGetAttributes says host memory (cudaHostAlloc'd or cudaHostRegister'ed)
bool st_is_pinned(const void *p);

// raise-only cudaFuncAttributeMaxDynamicSharedMemorySize (per device, per kernel): a
// concurrent caller with smaller tables can never lower the limit under a launch
int st_raise_smem_impl(const void *kern, int device, int bytes);
template <typename K>
static inline int st_raise_smem(K kern, int device, int bytes) {
    return bytes <= 48 * 1024 ? ST_OK : st_raise_smem_impl(reinterpret_cast<const void *>(kern), device, bytes);
}
