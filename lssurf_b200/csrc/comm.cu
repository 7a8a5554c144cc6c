// NCCL communicator wrapper: one process per GPU, all-reduce over NVLink 5 / NVSwitch.
//
// NCCL is bound at run time with dlopen so that the library shares the libnccl instance PyTorch
// already loaded (torch-bundled 2.28.x) instead of pulling in a second copy; only the stable C
// entry points are used.  The reference has no communication layer at all.
#include <dlfcn.h>

#include "common.cuh"

namespace lss {

typedef int ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef void *ncclComm_t;
enum { kNcclInt64 = 4, kNcclUint64 = 5, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

struct NcclApi {
    bool tried = false, ok = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.tried) return g_nccl.ok ? LSS_OK : fail(LSS_ENCCL, "libnccl.so.2 could not be loaded");
    g_nccl.tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(LSS_ENCCL, "dlopen(libnccl.so.2): %s", dlerror());
#define LOAD(field, name)                                                    \
    *(void **)(&g_nccl.field) = dlsym(h, name);                              \
    if (!g_nccl.field) return fail(LSS_ENCCL, "dlsym(%s) failed", name);
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.ok = true;
    return LSS_OK;
}

#define LSS_NCCL(call)                                                                            \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != 0) return fail(LSS_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
    } while (0)

int comm_allreduce(lss_comm *c, void *buf, int64_t n, int dtype, int op, cudaStream_t stream) {
    if (!c || c->n_ranks <= 1 || n <= 0) return LSS_OK;
    const int dt = dtype == 0 ? kNcclFloat64 : (dtype == 1 ? kNcclInt64 : kNcclUint64);
    const int ro = op == 0 ? kNcclSum : (op == 1 ? kNcclMin : kNcclMax);
    LSS_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, dt, ro, (ncclComm_t)c->nccl, stream));
    return LSS_OK;
}

}  // namespace lss

using namespace lss;

extern "C" int lss_comm_unique_id(char *id_h) {
    LSS_TRY(load_nccl());
    ncclUniqueId id;
    LSS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_h, id.internal, LSS_UNIQUE_ID_BYTES);
    return LSS_OK;
}

extern "C" int lss_comm_create(lss_comm_t **comm, const char *id_h, int32_t rank, int32_t n_ranks) {
    if (!comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(LSS_EINVAL, "comm_create: bad rank/n_ranks");
    lss_comm *c = new lss_comm();
    c->nccl = nullptr; c->rank = rank; c->n_ranks = n_ranks;
    if (n_ranks > 1) {
        int rc = load_nccl();
        if (rc != LSS_OK) { delete c; return rc; }
        ncclUniqueId id;
        memcpy(id.internal, id_h, LSS_UNIQUE_ID_BYTES);
        ncclComm_t nc = nullptr;
        ncclResult_t r = g_nccl.CommInitRank(&nc, n_ranks, id, rank);
        if (r != 0) { delete c; return fail(LSS_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
        c->nccl = nc;
    }
    *comm = c;
    return LSS_OK;
}

// ---- peer-to-peer mailbox over CUDA IPC ---------------------------------------------------------------------------
extern "C" int lss_comm_box_alloc(lss_comm_t *comm, int64_t bytes, char *handle_h) {
    if (!comm || bytes <= 0 || !handle_h) return fail(LSS_EINVAL, "comm_box_alloc: bad arguments");
    if (comm->n_ranks > LSS_MAX_RANKS) return fail(LSS_EINVAL, "comm_box_alloc: at most %d ranks", LSS_MAX_RANKS);
    if (comm->box_local) return fail(LSS_EINVAL, "comm_box_alloc: the mailbox exists already");
    // plain cudaMalloc (not the stream-ordered pool): IPC handles need a dedicated allocation
    LSS_CUDA(cudaMalloc(&comm->box_local, (size_t)bytes));
    LSS_CUDA(cudaMemset(comm->box_local, 0, (size_t)bytes));
    LSS_CUDA(cudaDeviceSynchronize());
    comm->box_bytes = (size_t)bytes;
    cudaIpcMemHandle_t h;
    memset(handle_h, 0, LSS_IPC_HANDLE_BYTES);
    if (comm->n_ranks > 1) {
        LSS_CUDA(cudaIpcGetMemHandle(&h, comm->box_local));
        static_assert(sizeof(h) <= LSS_IPC_HANDLE_BYTES, "IPC handle size");
        memcpy(handle_h, &h, sizeof(h));
    }
    return LSS_OK;
}

extern "C" int lss_comm_box_open(lss_comm_t *comm, const char *handles_h) {
    if (!comm || !comm->box_local) return fail(LSS_EINVAL, "comm_box_open: call lss_comm_box_alloc first");
    for (int r = 0; r < comm->n_ranks; ++r) {
        if (r == comm->rank) { comm->box_peer[r] = comm->box_local; continue; }
        if (!handles_h) return fail(LSS_EINVAL, "comm_box_open: null handles");
        cudaIpcMemHandle_t h;
        memcpy(&h, handles_h + (size_t)r * LSS_IPC_HANDLE_BYTES, sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&comm->box_peer[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(LSS_ECUDA, "comm_box_open: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
    }
    comm->box_open = true;
    return LSS_OK;
}

extern "C" int lss_comm_box_bytes(lss_comm_t *comm, int64_t *bytes_h) {
    if (!bytes_h) return fail(LSS_EINVAL, "comm_box_bytes: null output");
    *bytes_h = comm && comm->box_open ? (int64_t)comm->box_bytes : 0;
    return LSS_OK;
}

extern "C" int lss_comm_destroy(lss_comm_t *comm) {
    if (!comm) return LSS_OK;
    if (comm->box_local) {
        cudaDeviceSynchronize();
        for (int r = 0; r < comm->n_ranks; ++r)
            if (r != comm->rank && comm->box_peer[r]) cudaIpcCloseMemHandle(comm->box_peer[r]);
        cudaFree(comm->box_local);
        comm->box_local = nullptr;
    }
    if (comm->nccl && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)comm->nccl);
    delete comm;
    return LSS_OK;
}

extern "C" int lss_comm_allreduce_f64(lss_comm_t *comm, double *buf, int64_t n, void *stream) {
    return comm_allreduce(comm, buf, n, 0, 0, (cudaStream_t)stream);
}
