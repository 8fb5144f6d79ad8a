// fbgnn_comm.cu -- the one collective of the path: a sum of the int64 Monte-Carlo
// counters {frames, flagged, block errors, stage-0 failures} over the ranks of one box, by NCCL over
// NVLink / NVSwitch, on the context's own stream.  The reference has nothing here (one process per
// --gpu_id, n1270.py:10-26); frames shard by global frame id and no data crosses GPUs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) when the first communicator is created, so a
// single-GPU user needs no NCCL at all and the library has no link-time dependency on it.  Payloads
// live in a small device buffer owned by the context; the host arrays of the C ABI are staged through it.
#include <dlfcn.h>
#include <nccl.h>

#include "fbgnn_internal.h"

namespace {

struct NcclApi {
    void *dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.dl) return 0;
    const char *names[] = {getenv("FBGNN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *dl = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        dl = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (dl) break;
    }
    if (!dl) return fail(FBGNN_E_UNSUPPORTED, "libnccl.so.2 not found (%s); set FBGNN_NCCL_LIB", dlerror());
    NcclApi a;
    a.dl = dl;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(dl, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(dl, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(dl, "ncclAllReduce");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(dl, "ncclCommDestroy");
    a.GetVersion = (decltype(a.GetVersion))dlsym(dl, "ncclGetVersion");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(dl, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString) {
        dlclose(dl);
        return fail(FBGNN_E_UNSUPPORTED, "the NCCL library lacks a required symbol");
    }
    g_nccl = a;
    return 0;
}

}  // namespace

#define NCK(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(FBGNN_E_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_),   \
                        __FILE__, __LINE__);                                                       \
    } while (0)

extern "C" int fbgnn_comm_unique_id(uint8_t id[FBGNN_COMM_ID_BYTES]) {
    REQUIRE(id, "id is NULL");
    static_assert(FBGNN_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId u;
    NCK(g_nccl.GetUniqueId(&u));
    std::memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

extern "C" int fbgnn_comm_init_rank(fbgnn_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t id[FBGNN_COMM_ID_BYTES]) {
    REQUIRE(ctx && id, "NULL argument");
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    REQUIRE(!ctx->comm, "the context already has a communicator");
    if (int rc = nccl_load()) return rc;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    ncclUniqueId u;
    std::memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    NCK(g_nccl.CommInitRank(&comm, nranks, u, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    if (!ctx->comm_buf) CK(cudaMalloc(&ctx->comm_buf, FBGNN_COMM_MAX_ELEMS * 8));
    return 0;
}

extern "C" int fbgnn_comm_info(fbgnn_ctx *ctx, int32_t *nranks, int32_t *rank, int32_t *nccl_version) {
    REQUIRE(ctx, "ctx is NULL");
    if (nranks) *nranks = ctx->comm ? ctx->comm_size : 1;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (nccl_version) {
        int v = 0;
        if (g_nccl.dl && g_nccl.GetVersion) g_nccl.GetVersion(&v);
        *nccl_version = v;
    }
    return 0;
}

// One all-reduce of `count` 8-byte elements; host array in, host array out.  Without a communicator the
// call is the identity of a single-rank job (the values are already global).
static int allreduce8(fbgnn_ctx *ctx, void *host, int32_t count, ncclDataType_t dt, ncclRedOp_t op) {
    REQUIRE(ctx && host, "NULL argument");
    REQUIRE(count >= 0 && count <= FBGNN_COMM_MAX_ELEMS, "count %d out of range (max %d)", count, FBGNN_COMM_MAX_ELEMS);
    if (!ctx->comm || count == 0) return 0;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaMemcpyAsync(ctx->comm_buf, host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    NCK(g_nccl.AllReduce(ctx->comm_buf, ctx->comm_buf, (size_t)count, dt, op, (ncclComm_t)ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(host, ctx->comm_buf, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->collectives++;
    return 0;
}

extern "C" int fbgnn_allreduce_counters(fbgnn_ctx *ctx, int64_t *counters, int32_t count) {
    return allreduce8(ctx, counters, count, ncclInt64, ncclSum);
}

extern "C" int fbgnn_allreduce_f64(fbgnn_ctx *ctx, double *values, int32_t count, int32_t op) {
    REQUIRE(op == FBGNN_RED_SUM || op == FBGNN_RED_MAX, "unknown reduction %d", op);
    return allreduce8(ctx, values, count, ncclFloat64, op == FBGNN_RED_MAX ? ncclMax : ncclSum);
}

extern "C" int fbgnn_comm_barrier(fbgnn_ctx *ctx) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->comm) return 0;
    int64_t one = 1;
    return allreduce8(ctx, &one, 1, ncclInt64, ncclSum);
}

extern "C" int fbgnn_comm_destroy(fbgnn_ctx *ctx) {
    REQUIRE(ctx, "ctx is NULL");
    if (!ctx->comm) return 0;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    cudaStreamSynchronize(ctx->stream);
    ncclResult_t r = g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_size = 1;
    ctx->comm_rank = 0;
    if (r != ncclSuccess) return fail(FBGNN_E_CUDA, "ncclCommDestroy failed: %s", g_nccl.GetErrorString(r));
    return 0;
}
