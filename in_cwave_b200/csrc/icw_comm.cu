// icw_comm.cu -- the one place the path talks between GPUs: a time-sharded stream hands the
// state of its four half-band filters to the right neighbour, and the shards' clip counters / peaks are reduced.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy already in the process when torch loaded one):
// the library itself has no link-time dependency on NCCL, so it loads on boxes and in test containers without it,
// and every entry point below fails loudly with ICW_E_UNSUPPORTED if the symbols cannot be found.
// The communicator is ours (ncclCommInitRank from a 128-byte id the caller distributes however it likes: bench.py
// and in_cwave_b200/dist.py broadcast it with torch.distributed), or any ncclComm_t the caller already has.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>

#include <cuda_runtime.h>
#include <nccl.h>

#include "icw_internal.h"
#include "icw_comm.h"

namespace icw {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string why;
};

static NcclApi g_nccl;

const char *nccl_why() { return g_nccl.why.c_str(); }

bool nccl_load()
{
    NcclApi &n = g_nccl;
    if (n.h) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) {                      // the copy torch (or anybody) already mapped, if any
        n.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (n.h) break;
    }
    if (!n.h) {
        const char *env = getenv("ICW_NCCL_LIB");
        if (env && *env) n.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char *nm : names) {
        if (n.h) break;
        n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!n.h) { n.why = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?"); return false; }
#define ICW_SYM(field, name) \
    *(void **)(&n.field) = dlsym(n.h, name); \
    if (!n.field) { n.why = std::string("NCCL symbol missing: ") + name; n.h = nullptr; return false; }
    ICW_SYM(GetUniqueId, "ncclGetUniqueId");
    ICW_SYM(CommInitRank, "ncclCommInitRank");
    ICW_SYM(CommDestroy, "ncclCommDestroy");
    ICW_SYM(Send, "ncclSend");
    ICW_SYM(Recv, "ncclRecv");
    ICW_SYM(AllReduce, "ncclAllReduce");
    ICW_SYM(GroupStart, "ncclGroupStart");
    ICW_SYM(GroupEnd, "ncclGroupEnd");
    ICW_SYM(GetErrorString, "ncclGetErrorString");
    ICW_SYM(GetVersion, "ncclGetVersion");
#undef ICW_SYM
    return true;
}

int nccl_version()
{
    int v = 0;
    if (!nccl_load() || g_nccl.GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

#define NK(call)                                                                               \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) { err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return -2; } \
    } while (0)

int comm_unique_id(unsigned char out[ICW_COMM_ID_BYTES], std::string &err)
{
    if (!nccl_load()) { err = g_nccl.why; return -3; }
    static_assert(ICW_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, ICW_COMM_ID_BYTES);
    return 0;
}

int comm_init(const unsigned char idb[ICW_COMM_ID_BYTES], int rank, int world, void **comm, std::string &err)
{
    if (!nccl_load()) { err = g_nccl.why; return -3; }
    ncclUniqueId id;
    memcpy(id.internal, idb, ICW_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    NK(g_nccl.CommInitRank(&c, world, id, rank));
    *comm = (void *)c;
    return 0;
}

int comm_destroy(void *comm, std::string &err)
{
    if (!comm) return 0;
    if (!nccl_load()) { err = g_nccl.why; return -3; }
    NK(g_nccl.CommDestroy((ncclComm_t)comm));
    return 0;
}

// ring shift to the right: d_send (or NULL on the last rank) -> rank + 1, rank - 1 -> d_recv (or NULL on rank 0)
int comm_shift_right(void *comm, int rank, int world, const double *d_send, double *d_recv, size_t n_doubles,
                     cudaStream_t st, std::string &err)
{
    if (!nccl_load()) { err = g_nccl.why; return -3; }
    ncclComm_t c = (ncclComm_t)comm;
    NK(g_nccl.GroupStart());
    if (rank + 1 < world && d_send) NK(g_nccl.Send(d_send, n_doubles, ncclDouble, rank + 1, c, st));
    if (rank > 0 && d_recv) NK(g_nccl.Recv(d_recv, n_doubles, ncclDouble, rank - 1, c, st));
    NK(g_nccl.GroupEnd());
    return 0;
}

// d_sum[n_sum] (uint64) summed, d_max[n_max] (double) maximised over the ranks, in place
int comm_reduce(void *comm, unsigned long long *d_sum, size_t n_sum, double *d_max, size_t n_max, cudaStream_t st, std::string &err)
{
    if (!nccl_load()) { err = g_nccl.why; return -3; }
    ncclComm_t c = (ncclComm_t)comm;
    NK(g_nccl.GroupStart());
    NK(g_nccl.AllReduce(d_sum, d_sum, n_sum, ncclUint64, ncclSum, c, st));
    NK(g_nccl.AllReduce(d_max, d_max, n_max, ncclDouble, ncclMax, c, st));
    NK(g_nccl.GroupEnd());
    return 0;
}

}  // namespace icw
