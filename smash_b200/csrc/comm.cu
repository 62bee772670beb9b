// comm.cu -- the one collective of the path: a small all-reduce / all-gather over NCCL (NVLink / NVSwitch), one process per
// GPU, for the shared hyper-parameter gradient of the regionalised multi-catchment calibration and the
// gathered costs of a sharded ensemble.  libnccl.so.2 is opened at run time (no link-time dependency, no PyTorch); the
// communicator is built from a 128-byte unique id that rank 0 creates and the caller distributes (file or environment).
// Payloads are a few hundred values: they are staged through a small device buffer owned by the communicator, on a stream
// of its own.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/smash_b200.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;       // nccl.h:37-38
typedef void *NcclComm;
typedef int NcclResult;
enum { NCCL_INT32 = 2, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };  // nccl.h:280-286

struct Nccl {
    void *so = nullptr;
    NcclResult (*GetUniqueId)(NcclUniqueId *) = nullptr;
    NcclResult (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    NcclResult (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    NcclResult (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(NcclResult) = nullptr;
};

thread_local std::string g_comm_err;
int cfail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_comm_err = buf;
    return code;
}

Nccl *nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n.so ? &n : nullptr;
    tried = true;
    const char *names[] = {getenv("SMASH_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        n.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.so) break;
    }
    if (!n.so) return nullptr;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.so, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.so, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.so, "ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.so, "ncclAllReduce"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(n.so, "ncclAllGather"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.so, "ncclGetErrorString"));
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.AllGather || !n.GetErrorString) {
        dlclose(n.so);
        n.so = nullptr;
        return nullptr;
    }
    return &n;
}

}  // namespace

struct SmashComm {
    NcclComm comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    void *dbuf = nullptr;
    size_t dbytes = 0;
    int ensure(size_t bytes) {
        if (bytes <= dbytes) return 0;
        if (dbuf) cudaFree(dbuf);
        dbuf = nullptr; dbytes = 0;
        if (cudaMalloc(&dbuf, bytes) != cudaSuccess) { cudaGetLastError(); return 1; }
        dbytes = bytes;
        return 0;
    }
};

extern "C" const char *smash_b200_comm_last_error(void) { return g_comm_err.c_str(); }

extern "C" int smash_b200_comm_unique_id(char id[128]) {
    Nccl *n = nccl();
    if (!n) return cfail(SMASH_B200_ENODEV, "libnccl.so.2 could not be loaded");
    NcclUniqueId u;
    const NcclResult r = n->GetUniqueId(&u);
    if (r != 0) return cfail(SMASH_B200_ECUDA, "ncclGetUniqueId: %s", n->GetErrorString(r));
    memcpy(id, u.internal, 128);
    return 0;
}

extern "C" int smash_b200_comm_create(const char id[128], int32_t rank, int32_t world, SmashComm **out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return cfail(SMASH_B200_EINVAL, "comm_create: bad arguments");
    Nccl *n = nccl();
    if (!n) return cfail(SMASH_B200_ENODEV, "libnccl.so.2 could not be loaded");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return cfail(SMASH_B200_ENODEV, "no CUDA device"); }
    SmashComm *c = new SmashComm();
    c->rank = rank; c->world = world;
    NcclUniqueId u;
    memcpy(u.internal, id, 128);
    const NcclResult r = n->CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; return cfail(SMASH_B200_ECUDA, "ncclCommInitRank: %s", n->GetErrorString(r)); }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        n->CommDestroy(c->comm); delete c;
        return cfail(SMASH_B200_ECUDA, "cudaStreamCreate failed");
    }
    *out = c;
    return 0;
}

extern "C" void smash_b200_comm_destroy(SmashComm *c) {
    if (!c) return;
    Nccl *n = nccl();
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->dbuf) cudaFree(c->dbuf);
    if (n && c->comm) n->CommDestroy(c->comm);
    delete c;
}

static int dtype_of(int32_t kind, size_t *size) {
    switch (kind) {
        case 0: *size = 4; return NCCL_FLOAT32;
        case 1: *size = 8; return NCCL_FLOAT64;
        default: *size = 4; return NCCL_INT32;
    }
}

// in-place all-reduce of a host vector: kind 0 float32, 1 float64, 2 int32; op 0 sum, 2 max, 3 min (NCCL's codes)
extern "C" int smash_b200_comm_allreduce(SmashComm *c, void *host, int64_t count, int32_t kind, int32_t op) {
    if (!c || !host || count < 0) return cfail(SMASH_B200_EINVAL, "comm_allreduce: bad arguments");
    if (count == 0) return 0;
    Nccl *n = nccl();
    size_t es;
    const int dt = dtype_of(kind, &es);
    if (c->ensure((size_t)count * es)) return cfail(SMASH_B200_ENOMEM, "comm_allreduce: device buffer");
    if (cudaMemcpyAsync(c->dbuf, host, (size_t)count * es, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return cfail(SMASH_B200_ECUDA, "comm_allreduce: upload failed");
    const NcclResult r = n->AllReduce(c->dbuf, c->dbuf, (size_t)count, dt, op, c->comm, c->stream);
    if (r != 0) return cfail(SMASH_B200_ECUDA, "ncclAllReduce: %s", n->GetErrorString(r));
    if (cudaMemcpyAsync(host, c->dbuf, (size_t)count * es, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
        return cfail(SMASH_B200_ECUDA, "comm_allreduce: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// all-gather: every rank contributes `count` values from `send`, `recv` receives world * count values in rank order
extern "C" int smash_b200_comm_allgather(SmashComm *c, const void *send, void *recv, int64_t count, int32_t kind) {
    if (!c || !send || !recv || count < 0) return cfail(SMASH_B200_EINVAL, "comm_allgather: bad arguments");
    if (count == 0) return 0;
    Nccl *n = nccl();
    size_t es;
    const int dt = dtype_of(kind, &es);
    const size_t part = (size_t)count * es;
    if (c->ensure(part * (size_t)(c->world + 1))) return cfail(SMASH_B200_ENOMEM, "comm_allgather: device buffer");
    char *d = static_cast<char *>(c->dbuf);
    if (cudaMemcpyAsync(d, send, part, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return cfail(SMASH_B200_ECUDA, "comm_allgather: upload failed");
    const NcclResult r = n->AllGather(d, d + part, (size_t)count, dt, c->comm, c->stream);
    if (r != 0) return cfail(SMASH_B200_ECUDA, "ncclAllGather: %s", n->GetErrorString(r));
    if (cudaMemcpyAsync(recv, d + part, part * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
        return cfail(SMASH_B200_ECUDA, "comm_allgather: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

extern "C" int smash_b200_comm_rank(const SmashComm *c) { return c ? c->rank : -1; }
extern "C" int smash_b200_comm_world(const SmashComm *c) { return c ? c->world : -1; }
