// NCCL plumbing of a sharded colony (wr_acs_comm_init).  One process per GPU; NCCL is used for what it is
// good at here — the rendezvous and the one-off exchange of 64-byte CUDA IPC handles when a search begins — while the
// per-iteration exchange runs over NVLink peer memory inside our own kernels (acs.cu, k_peer_barrier).
//
// libnccl.so.2 is opened at run time (dlopen): libwrgpu.so keeps no link-time dependency on it, a process that already
// carries an NCCL (torch's bundled one) shares it, and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "wr_internal.cuh"

namespace wr {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return WR_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (e.g. torch's)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("wr_acs_comm_init: cannot load libnccl.so.2: %s", dlerror()); return WR_ERR_CUDA; }
    NcclApi api;
    api.lib = h;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
        set_error("wr_acs_comm_init: libnccl.so.2 lacks a required symbol");
        return WR_ERR_CUDA;
    }
    g_nccl = api;
    return WR_OK;
}

#define WR_NCCL(expr)                                                                              \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            wr::set_error("%s failed: %s (%s:%d)", #expr, wr::g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
            return WR_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// one communicator per process: searches created in a loop (one per request) share it
struct CommCache {
    unsigned char id[WR_COMM_ID_BYTES] = {};
    int rank = -1, nranks = 0, device = -1;
    ncclComm_t comm = nullptr;
};
static CommCache g_comm;

int comm_get(const void* unique_id, int rank, int nranks, void** comm_out)
{
    int rc = nccl_load();
    if (rc != WR_OK) return rc;
    int dev = 0;
    WR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_comm.comm && g_comm.rank == rank && g_comm.nranks == nranks && g_comm.device == dev && memcmp(g_comm.id, unique_id, WR_COMM_ID_BYTES) == 0) {
        *comm_out = g_comm.comm;
        return WR_OK;
    }
    if (g_comm.comm) { g_nccl.CommDestroy(g_comm.comm); g_comm = CommCache(); }
    static_assert(sizeof(ncclUniqueId) == WR_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclComm_t c = nullptr;
    WR_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
    memcpy(g_comm.id, unique_id, WR_COMM_ID_BYTES);
    g_comm.rank = rank; g_comm.nranks = nranks; g_comm.device = dev; g_comm.comm = c;
    *comm_out = c;
    return WR_OK;
}

// all_gather of `bytes` per rank; device buffers; enqueued on `s`
int comm_all_gather(void* comm, const void* d_send, void* d_recv, size_t bytes, cudaStream_t s)
{
    WR_NCCL(g_nccl.AllGather(d_send, d_recv, bytes, ncclUint8, static_cast<ncclComm_t>(comm), s));
    return WR_OK;
}

void comm_release()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_comm.comm) { g_nccl.CommDestroy(g_comm.comm); g_comm = CommCache(); }
}

}  // namespace wr

extern "C" int wr_comm_unique_id(void* id)
{
    WR_REQUIRE(id, WR_ERR_INVALID, "wr_comm_unique_id: null");
    int rc = wr::nccl_load();
    if (rc != WR_OK) return rc;
    ncclUniqueId u;
    WR_NCCL(wr::g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return WR_OK;
}
