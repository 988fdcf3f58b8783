// comm.cu — NCCL transport of the slab exchange: one fixed-size message to rank-1 and one to rank+1
// per step, ncclSend/ncclRecv inside one group on the library's stream.  libnccl is opened at run time (dlopen), so a
// single-GPU user needs no NCCL and a process that already loaded torch's bundled libnccl.so.2 shares that copy.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "t2d_internal.h"

namespace t2d {

// the few NCCL declarations needed (nccl.h, NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0 } ncclDataType_t;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static bool nccl_load(std::string* err)
{
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        *err = std::string("cannot open libnccl.so.2: ") + dlerror();
        return false;
    }
    NcclApi a;
    a.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) {
            ok = false;
            *err = std::string("libnccl lacks ") + name;
        }
        return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return false;
    g_nccl = a;
    return true;
}

struct NcclLink {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

#define NCK(call)                                                                               \
    do {                                                                                        \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess) {                                                               \
            *err = std::string(#call " failed: ") + g_nccl.GetErrorString(r__);                 \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

int nccl_unique_id(uint8_t* out, std::string* err)
{
    if (!nccl_load(err)) return -1;
    static_assert(sizeof(ncclUniqueId) == T2D_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

NcclLink* nccl_link_create(int rank, int world, const uint8_t* id_bytes, std::string* err)
{
    if (!nccl_load(err)) return nullptr;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    NcclLink* l = new NcclLink();
    l->rank = rank;
    l->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&l->comm, world, id, rank);
    if (r != ncclSuccess) {
        *err = std::string("ncclCommInitRank failed: ") + g_nccl.GetErrorString(r);
        delete l;
        return nullptr;
    }
    return l;
}

void nccl_link_destroy(NcclLink* l)
{
    if (!l) return;
    if (l->comm) g_nccl.CommDestroy(l->comm);
    delete l;
}

// one message of `bytes` to each existing neighbour and one from each, in a single NCCL group on stream s
// + the far message (far_bytes; far_recv has one slot per rank) to and from every other rank
int nccl_exchange(NcclLink* l, const void* send_left, void* recv_left, const void* send_right, void* recv_right, size_t bytes,
                  const void* far_send, void* far_recv, size_t far_bytes, cudaStream_t s, std::string* err)
{
    NCK(g_nccl.GroupStart());
    if (l->world > 1 && far_bytes) {
        for (int r = 0; r < l->world; ++r) {
            if (r == l->rank) continue;
            NCK(g_nccl.Send(far_send, far_bytes, ncclChar, r, l->comm, s));
            NCK(g_nccl.Recv(static_cast<char*>(far_recv) + (size_t)r * far_bytes, far_bytes, ncclChar, r, l->comm, s));
        }
    }
    if (l->rank > 0) {
        NCK(g_nccl.Send(send_left, bytes, ncclChar, l->rank - 1, l->comm, s));
        NCK(g_nccl.Recv(recv_left, bytes, ncclChar, l->rank - 1, l->comm, s));
    }
    if (l->rank < l->world - 1) {
        NCK(g_nccl.Send(send_right, bytes, ncclChar, l->rank + 1, l->comm, s));
        NCK(g_nccl.Recv(recv_right, bytes, ncclChar, l->rank + 1, l->comm, s));
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

}  // namespace t2d
