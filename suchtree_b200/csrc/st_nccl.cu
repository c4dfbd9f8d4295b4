// The path's ONE collective: the sum of six fp64 moments across the ranks of a box
// (NCCL over NVLink 5 / NVSwitch), enqueued on the stream of the kernel that produced
// them.  libnccl is resolved at run time (dlopen): a single-GPU process never needs it,
// and inside a torch process the copy torch already loaded is the one that is found.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "st_internal.cuh"

namespace {

// the few NCCL declarations used (stable ABI since NCCL 2.0; nccl.h:37-38, 146, 160, 181, 215, 260, 286, 392)
typedef struct { char internal[128]; } nccl_unique_id;
typedef void *nccl_comm_t;
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    bool tried = false, ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool load_nccl() {
    std::lock_guard<std::mutex> l(g_nccl_mu);
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    const char *names[] = {getenv("SUCHTREE_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !nm[0]) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        st_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
        return false;
    }
#define ST_SYM(field, name)                                                          \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name)); \
    if (!g_nccl.field) {                                                             \
        st_set_error("NCCL symbol %s missing", name);                                \
        return false;                                                                \
    }
    ST_SYM(GetUniqueId, "ncclGetUniqueId");
    ST_SYM(CommInitRank, "ncclCommInitRank");
    ST_SYM(CommDestroy, "ncclCommDestroy");
    ST_SYM(AllReduce, "ncclAllReduce");
    ST_SYM(GetErrorString, "ncclGetErrorString");
    ST_SYM(GetVersion, "ncclGetVersion");
#undef ST_SYM
    g_nccl.ok = true;
    return true;
}

int nccl_fail(const char *what, int rc) {
    st_set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return ST_ERR_CUDA;
}

}  // namespace

extern "C" int st_nccl_version(int *version) {
    if (!version) return ST_ERR_INVALID_ARG;
    if (!load_nccl()) return ST_ERR_CUDA;
    const int rc = g_nccl.GetVersion(version);
    return rc == NCCL_SUCCESS ? ST_OK : nccl_fail("ncclGetVersion", rc);
}

extern "C" int st_nccl_unique_id(void *id128) {
    if (!id128) return ST_ERR_INVALID_ARG;
    if (!load_nccl()) return ST_ERR_CUDA;
    nccl_unique_id id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != NCCL_SUCCESS) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128, id.internal, sizeof(id.internal));
    return ST_OK;
}

extern "C" int st_nccl_comm_create(int device, int world, int rank, const void *id128, void **comm) {
    if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) {
        st_set_error("st_nccl_comm_create: bad arguments");
        return ST_ERR_INVALID_ARG;
    }
    *comm = nullptr;
    if (!load_nccl()) return ST_ERR_CUDA;
    DeviceGuard g(device);
    if (!g.ok) {
        st_set_error("st_nccl_comm_create: cudaSetDevice(%d) failed", device);
        return ST_ERR_CUDA;
    }
    nccl_unique_id id;
    memcpy(id.internal, id128, sizeof(id.internal));
    nccl_comm_t c = nullptr;
    const int rc = g_nccl.CommInitRank(&c, world, id, rank);
    if (rc != NCCL_SUCCESS) return nccl_fail("ncclCommInitRank", rc);
    *comm = c;
    return ST_OK;
}

extern "C" int st_nccl_comm_destroy(void *comm) {
    if (!comm) return ST_OK;
    if (!load_nccl()) return ST_ERR_CUDA;
    const int rc = g_nccl.CommDestroy(static_cast<nccl_comm_t>(comm));
    return rc == NCCL_SUCCESS ? ST_OK : nccl_fail("ncclCommDestroy", rc);
}

// in-place sum of `count` doubles at d_buf over the communicator, asynchronous on `stream`
int st_nccl_allreduce_sum_f64(void *comm, double *d_buf, int count, cudaStream_t stream) {
    if (!load_nccl()) return ST_ERR_CUDA;
    const int rc = g_nccl.AllReduce(d_buf, d_buf, size_t(count), NCCL_FLOAT64, NCCL_SUM,
                                    static_cast<nccl_comm_t>(comm), stream);
    return rc == NCCL_SUCCESS ? ST_OK : nccl_fail("ncclAllReduce", rc);
}
