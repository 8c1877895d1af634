// NCCL plumbing for the two row-sharded paths (TSQR R all-gather, Gram all-reduce).
// One process per GPU: rank 0 creates the unique id, the host side (Python, any channel) hands it
// to every rank, each rank calls lq_comm_init on its own context.  libnccl is dlopen'ed lazily so
// the library loads (and every single-GPU path works) on machines without NCCL.
#include <dlfcn.h>

#include <mutex>

#include "../../include/linalg_b200.h"
#include "ops.cuh"

namespace lq {
namespace {

struct NcclUniqueId {
    char internal[128];
};
using ncclComm_t = void*;
constexpr int kNcclFloat64 = 8;
constexpr int kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("LINALG_B200_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nme : names) {
            if (!nme) continue;
            a.handle = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
            a.why = dlerror();
        }
        if (!a.handle) return;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.AllGather) {
            a.why = "libnccl is missing required symbols";
            a.handle = nullptr;
        }
    });
    return a;
}

int nccl_fail(Ctx* c, const char* what, int rc) {
    NcclApi& a = api();
    set_error(c, "%s: NCCL error %d (%s)", what, rc, a.GetErrorString ? a.GetErrorString(rc) : "?");
    return LQ_ERR_NCCL_BASE + rc;
}

}  // namespace

int comm_allreduce_sum(Ctx* c, double* buf, long long count) {
    if (c->nranks <= 1) return LQ_OK;
    LQ_REQUIRE(c, c->nccl_comm != nullptr, LQ_ERR_ARG, "communicator not initialised (lq_comm_init)");
    int rc = api().AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream);
    if (rc != 0) return nccl_fail(c, "ncclAllReduce", rc);
    return LQ_OK;
}

int comm_allgather(Ctx* c, const double* send, double* recv, long long count_per_rank) {
    if (c->nranks <= 1) {
        if (send != recv)
            LQ_CUDA(c, cudaMemcpyAsync(recv, send, sizeof(double) * (size_t)count_per_rank, cudaMemcpyDeviceToDevice, c->stream));
        return LQ_OK;
    }
    LQ_REQUIRE(c, c->nccl_comm != nullptr, LQ_ERR_ARG, "communicator not initialised (lq_comm_init)");
    int rc = api().AllGather(send, recv, (size_t)count_per_rank, kNcclFloat64, c->nccl_comm, c->stream);
    if (rc != 0) return nccl_fail(c, "ncclAllGather", rc);
    return LQ_OK;
}

}  // namespace lq

using namespace lq;

extern "C" {

int lq_comm_unique_id(void* id128) {
    if (!id128) return LQ_ERR_ARG;
    NcclApi& a = api();
    if (!a.handle) {
        set_error(nullptr, "NCCL not available: %s", a.why.c_str());
        return LQ_ERR_NCCL_MISSING;
    }
    NcclUniqueId id;
    int rc = a.GetUniqueId(&id);
    if (rc != 0) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return LQ_OK;
}

int lq_comm_init(lq_ctx* h, int nranks, int rank, const void* id128) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, nranks >= 1 && rank >= 0 && rank < nranks, LQ_ERR_ARG, "bad rank %d of %d", rank, nranks);
    lq_comm_destroy(h);
    c->nranks = nranks;
    c->rank = rank;
    if (nranks == 1) return LQ_OK;
    LQ_REQUIRE(c, id128 != nullptr, LQ_ERR_ARG, "null unique id");
    NcclApi& a = api();
    if (!a.handle) {
        set_error(c, "NCCL not available: %s", a.why.c_str());
        return LQ_ERR_NCCL_MISSING;
    }
    LQ_CUDA(c, cudaSetDevice(c->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    int rc = a.CommInitRank(&comm, nranks, id, rank);
    if (rc != 0) return nccl_fail(c, "ncclCommInitRank", rc);
    c->nccl_comm = comm;
    return LQ_OK;
}

int lq_comm_destroy(lq_ctx* h) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_OK;
    if (c->nccl_comm) {
        cudaStreamSynchronize(c->stream);
        api().CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    c->nranks = 1;
    c->rank = 0;
    return LQ_OK;
}

int lq_comm_allreduce_sum(lq_ctx* h, double* dbuf, int64_t count) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_CUDA(c, cudaSetDevice(c->device));
    return comm_allreduce_sum(c, dbuf, count);
}

int lq_comm_allgather(lq_ctx* h, const double* dsend, double* drecv, int64_t count_per_rank) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_CUDA(c, cudaSetDevice(c->device));
    return comm_allgather(c, dsend, drecv, count_per_rank);
}

}  // extern "C"
