// fbgnn_api.cu -- C ABI (include/fbgnn.h) over the kernels in fbgnn_kernels.cuh.
// Host-side handle management, graph-table construction, launch configuration and the
// multi-launch orchestration of the fused Monte-Carlo pipelines.  No CPU compute path:
// every entry point that does work launches CUDA kernels and fails with FBGNN_E_CUDA when
// there is no device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fbgnn.h"
#include "fbgnn_kernels.cuh"
#include "fbgnn_gbp_tc.cuh"
#include "fbgnn_train.cuh"

using namespace fbgnn;

// ------------------------------------------------------------------ errors -------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(FBGNN_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

#define REQUIRE(cond, ...)                                   \
    do {                                                     \
        if (!(cond)) return fail(FBGNN_E_INVALID, __VA_ARGS__); \
    } while (0)

// ------------------------------------------------------------------ handles ------------
struct Workspace {                 // per-context scratch of the fused pipelines
    int64_t cap_frames = 0;
    int n = 0, m = 0;
    uint8_t *vbits = nullptr, *sbits = nullptr, *active[2] = {nullptr, nullptr}, *rounds = nullptr;
    float *L = nullptr, *P = nullptr, *logit = nullptr;
    int *list[2] = {nullptr, nullptr};
    int *list_count = nullptr;     // [2]
    unsigned long long *counters = nullptr;   // [4]
    void *hard = nullptr;          // binary pipeline: unused (decisions live in vbits)
};

struct fbgnn_ctx {
    int device = 0, num_sms = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    int math_mode = FBGNN_MATH_EXACT;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    char name[256] = {0};
    Workspace ws;
};

struct fbgnn_graph {
    fbgnn_ctx *ctx;
    SideDev dev;
    int max_dc = 0, max_dv = 0;
    bool decodable = true;         // false: only the bit-packed rows exist (dense logical matrices)
    std::string why_not;
    std::vector<void *> allocs;
};

struct fbgnn_code {
    fbgnn_ctx *ctx;
    fbgnn_graph *X, *Z;
    int kx = 0, kz = 0, W = 0;
    uint32_t *lx_bits = nullptr, *lz_bits = nullptr;
    // host copies of the CSR of hx / hz (to cut row bases out of them) and the OSD-0 bases
    std::vector<int32_t> hx_ptr, hx_idx, hz_ptr, hz_idx;
    int *lx_ptr = nullptr, *lz_ptr = nullptr;                // CSR rows of the logical operators (GNN_BP4 logits)
    idx_t *lx_col = nullptr, *lz_col = nullptr;
    fbgnn_graph *basis_x = nullptr, *basis_z = nullptr;      // hx[pivot_hx], hz[pivot_hz]
    idx_t *pivot_x = nullptr, *pivot_z = nullptr;            // device [rank]
};

struct fbgnn_gnn {
    fbgnn_ctx *ctx;
    int H, M, act, reduce, use_bias;
    float *weights = nullptr;      // packed GnnLayout<H,M>
    int total = 0;
};

static int set_device(fbgnn_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    return 0;
}

static int need_decodable(const fbgnn_graph *g) {
    if (!g->decodable) return fail(FBGNN_E_UNSUPPORTED, "%s", g->why_not.c_str());
    return 0;
}

template <typename T> static View2<T> v2(const fbgnn_tensor2 &t) { return View2<T>{(T *)t.ptr, t.s0, t.s1}; }
template <typename T> static View3<T> v3(const fbgnn_tensor3 &t) { return View3<T>{(T *)t.ptr, t.s0, t.s1, t.s2}; }

// ------------------------------------------------------------------ library / context ---
extern "C" int fbgnn_version(void) { return FBGNN_VERSION; }
extern "C" const char *fbgnn_last_error(void) { return g_err.c_str(); }

extern "C" int fbgnn_device_count(int *count) {
    REQUIRE(count, "count is NULL");
    CK(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int fbgnn_ctx_create(int device, fbgnn_ctx **out) {
    REQUIRE(out, "ctx is NULL");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    REQUIRE(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
    CK(cudaSetDevice(device));
    fbgnn_ctx *ctx = new fbgnn_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    snprintf(ctx->name, sizeof ctx->name, "%s", prop.name);
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CK(cudaEventCreate(&ctx->ev0));
    CK(cudaEventCreate(&ctx->ev1));
    *out = ctx;
    return 0;
}

static void ws_free(Workspace &w) {
    cudaFree(w.vbits); cudaFree(w.sbits); cudaFree(w.active[0]); cudaFree(w.active[1]);
    cudaFree(w.rounds); cudaFree(w.L); cudaFree(w.P); cudaFree(w.logit); cudaFree(w.list[0]);
    cudaFree(w.list[1]); cudaFree(w.list_count); cudaFree(w.counters);
    w = Workspace();
}

extern "C" int fbgnn_ctx_destroy(fbgnn_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ws_free(ctx->ws);
    cudaFree(ctx->flush_buf);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

extern "C" int fbgnn_ctx_sync(fbgnn_ctx *ctx) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int fbgnn_ctx_device(fbgnn_ctx *ctx, int *device, int *num_sms, char *name, int name_len) {
    REQUIRE(ctx, "ctx is NULL");
    if (device) *device = ctx->device;
    if (num_sms) *num_sms = ctx->num_sms;
    if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", ctx->name);
    return 0;
}

extern "C" int fbgnn_timer_start(fbgnn_ctx *ctx) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}

extern "C" int fbgnn_timer_stop(fbgnn_ctx *ctx, float *ms) {
    REQUIRE(ctx && ms, "NULL argument");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return 0;
}

extern "C" int fbgnn_ctx_set_math(fbgnn_ctx *ctx, int32_t mode) {
    REQUIRE(ctx, "ctx is NULL");
    REQUIRE(mode == FBGNN_MATH_EXACT || mode == FBGNN_MATH_FAST, "unknown math mode %d", mode);
    ctx->math_mode = mode;
    return 0;
}

extern "C" int fbgnn_ctx_get_math(fbgnn_ctx *ctx, int32_t *mode) {
    REQUIRE(ctx && mode, "NULL argument");
    *mode = ctx->math_mode;
    return 0;
}

extern "C" int fbgnn_launch_count(fbgnn_ctx *ctx, int64_t *launches) {
    REQUIRE(ctx && launches, "NULL argument");
    *launches = ctx->launches;
    return 0;
}

// ------------------------------------------------------------------ memory --------------
// Stream-ordered allocation from the device's default memory pool (kept cached: the release
// threshold is raised at context creation), so per-call output tensors cost no cudaMalloc.
extern "C" int fbgnn_malloc(fbgnn_ctx *ctx, size_t bytes, void **dptr) {
    REQUIRE(ctx && dptr, "NULL argument");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    cudaError_t e = cudaMallocAsync(dptr, bytes ? bytes : 1, ctx->stream);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(FBGNN_E_NOMEM, "cudaMallocAsync(%zu) out of memory", bytes); }
    CK(e);
    return 0;
}
extern "C" int fbgnn_free(fbgnn_ctx *ctx, void *dptr) {
    REQUIRE(ctx, "ctx is NULL");
    if (!dptr) return 0;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaFreeAsync(dptr, ctx->stream));
    return 0;
}
extern "C" int fbgnn_memset(fbgnn_ctx *ctx, void *dptr, int value, size_t bytes) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return 0;
}
extern "C" int fbgnn_memcpy_h2d(fbgnn_ctx *ctx, void *dst, const void *src, size_t bytes) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int fbgnn_memcpy_d2h(fbgnn_ctx *ctx, void *dst, const void *src, size_t bytes) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int fbgnn_memcpy_d2d(fbgnn_ctx *ctx, void *dst, const void *src, size_t bytes) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
extern "C" int fbgnn_host_alloc(size_t bytes, void **hptr) {
    REQUIRE(hptr, "hptr is NULL");
    CK(cudaMallocHost(hptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int fbgnn_host_free(void *hptr) {
    CK(cudaFreeHost(hptr));
    return 0;
}
extern "C" int fbgnn_flush_l2(fbgnn_ctx *ctx) {
    REQUIRE(ctx, "ctx is NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;        // 256 MiB > 126 MB of L2
        CK(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    k_fill<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((uint32_t *)ctx->flush_buf,
                                                      (int64_t)(ctx->flush_bytes / 4), 0u);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ graphs --------------
template <typename T>
static int upload(fbgnn_graph *g, const std::vector<T> &h, const T **dptr) {
    void *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    g->allocs.push_back(d);
    if (!h.empty()) CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dptr = (const T *)d;
    return 0;
}

static void pack_rows(int32_t n, int32_t m, const int32_t *indptr, const int32_t *indices,
                      std::vector<uint32_t> &bits) {
    const int W = (n + 31) / 32;
    bits.assign((size_t)std::max(m, 0) * W, 0u);
    for (int r = 0; r < m; r++)
        for (int k = indptr[r]; k < indptr[r + 1]; k++)
            bits[(size_t)r * W + (indices[k] >> 5)] ^= 1u << (indices[k] & 31);
}

static int validate_csr(int32_t n, int32_t m, const int32_t *indptr, const int32_t *indices, const char *what) {
    REQUIRE(n > 0 && m >= 0, "%s: bad shape (%d x %d)", what, m, n);
    REQUIRE(indptr && (indices || indptr[m] == 0), "%s: NULL CSR arrays", what);
    REQUIRE(indptr[0] == 0, "%s: indptr[0] != 0", what);
    for (int r = 0; r < m; r++) {
        REQUIRE(indptr[r + 1] >= indptr[r], "%s: indptr not monotone at row %d", what, r);
        for (int k = indptr[r]; k < indptr[r + 1]; k++) {
            REQUIRE(indices[k] >= 0 && indices[k] < n, "%s: column index %d out of range in row %d", what, indices[k], r);
            REQUIRE(k == indptr[r] || indices[k] > indices[k - 1], "%s: row %d not strictly increasing", what, r);
        }
    }
    return 0;
}

extern "C" int fbgnn_graph_create(fbgnn_ctx *ctx, int32_t n, int32_t m, const int32_t *indptr,
                                  const int32_t *indices, fbgnn_graph **out) {
    REQUIRE(ctx && out, "NULL argument");
    if (int rc = validate_csr(n, m, indptr, indices, "graph")) return rc;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const int E = indptr[m];
    fbgnn_graph *g = new fbgnn_graph();
    g->ctx = ctx;
    int max_dc = 0, max_dv = 0, min_dc = 1 << 30, min_dv = 1 << 30;
    for (int c = 0; c < m; c++) {
        const int dc = indptr[c + 1] - indptr[c];
        max_dc = std::max(max_dc, dc); min_dc = std::min(min_dc, dc);
    }
    char why[256] = {0};
    if (n > 65535 || m > 65535 || E > 65535)
        snprintf(why, sizeof why, "matrix too large for the shared-memory resident decoder (n=%d m=%d edges=%d; "
                 "limit 65535 each)", n, m, E);
    else if (max_dc > 64)
        snprintf(why, sizeof why, "check degree %d > 64 is not supported by the decoder", max_dc);
    g->decodable = why[0] == 0;
    g->why_not = why;
    // CN order is the CSR order (sorted by (cn, vn)); VN order = stable counting sort by vn.
    std::vector<idx_t> cn_ptr, cn_vn, cn_edge, vn_ptr, vn_cn;
    if (g->decodable) {
        cn_ptr.resize(m + 1); cn_vn.resize(E); cn_edge.resize(E); vn_ptr.resize(n + 1); vn_cn.resize(E);
        std::vector<int> deg(n, 0), fill(n, 0);
        for (int k = 0; k < E; k++) deg[indices[k]]++;
        int acc = 0;
        for (int v = 0; v < n; v++) { vn_ptr[v] = (idx_t)acc; fill[v] = acc; acc += deg[v]; }
        vn_ptr[n] = (idx_t)acc;
        for (int c = 0; c < m; c++) {
            cn_ptr[c] = (idx_t)indptr[c];
            for (int k = indptr[c]; k < indptr[c + 1]; k++) {
                const int v = indices[k];
                const int pos = fill[v]++;
                cn_vn[k] = (idx_t)v;
                cn_edge[k] = (idx_t)pos;
                vn_cn[pos] = (idx_t)c;
            }
        }
        cn_ptr[m] = (idx_t)E;
        for (int v = 0; v < n; v++) { max_dv = std::max(max_dv, deg[v]); min_dv = std::min(min_dv, deg[v]); }
    }
    g->max_dc = max_dc; g->max_dv = max_dv;
    std::vector<uint32_t> bits;
    pack_rows(n, m, indptr, indices, bits);
    SideDev &d = g->dev;
    d.n = n; d.m = m; d.E = E;
    d.reg_dc = (m > 0 && max_dc == min_dc) ? max_dc : 0;
    d.reg_dv = (max_dv == min_dv) ? max_dv : 0;
    int rc = 0;
    rc |= upload(g, vn_ptr, &d.vn_ptr); rc |= upload(g, vn_cn, &d.vn_cn);
    rc |= upload(g, cn_ptr, &d.cn_ptr); rc |= upload(g, cn_edge, &d.cn_edge);
    rc |= upload(g, cn_vn, &d.cn_vn);   rc |= upload(g, bits, &d.bitrows);
    if (rc) { fbgnn_graph_destroy(g); return FBGNN_E_CUDA; }
    *out = g;
    return 0;
}

extern "C" int fbgnn_graph_destroy(fbgnn_graph *g) {
    if (!g) return 0;
    cudaSetDevice(g->ctx->device);
    for (void *p : g->allocs) cudaFree(p);
    delete g;
    return 0;
}

extern "C" int fbgnn_code_create(fbgnn_ctx *ctx, int32_t n, int32_t m_x, const int32_t *hx_indptr,
                                 const int32_t *hx_indices, int32_t m_z, const int32_t *hz_indptr,
                                 const int32_t *hz_indices, int32_t k_x, const int32_t *lx_indptr,
                                 const int32_t *lx_indices, int32_t k_z, const int32_t *lz_indptr,
                                 const int32_t *lz_indices, fbgnn_code **out) {
    REQUIRE(ctx && out, "NULL argument");
    fbgnn_code *c = new fbgnn_code();
    c->ctx = ctx;
    c->X = c->Z = nullptr;
    int rc = fbgnn_graph_create(ctx, n, m_x, hx_indptr, hx_indices, &c->X);
    if (!rc) rc = fbgnn_graph_create(ctx, n, m_z, hz_indptr, hz_indices, &c->Z);
    if (!rc) rc = need_decodable(c->X);
    if (!rc) rc = need_decodable(c->Z);
    if (!rc && k_x > 0) rc = validate_csr(n, k_x, lx_indptr, lx_indices, "lx");
    if (!rc && k_z > 0) rc = validate_csr(n, k_z, lz_indptr, lz_indices, "lz");
    if (rc) { fbgnn_code_destroy(c); return rc; }
    c->kx = std::max(k_x, 0); c->kz = std::max(k_z, 0); c->W = (n + 31) / 32;
    c->hx_ptr.assign(hx_indptr, hx_indptr + m_x + 1); c->hx_idx.assign(hx_indices, hx_indices + hx_indptr[m_x]);
    c->hz_ptr.assign(hz_indptr, hz_indptr + m_z + 1); c->hz_idx.assign(hz_indices, hz_indices + hz_indptr[m_z]);
    std::vector<uint32_t> bits;
    auto up_rows = [&](int k, const int32_t *ptr, const int32_t *idx, int **dptr, idx_t **dcol) -> int {
        std::vector<int> p(ptr, ptr + k + 1);
        std::vector<idx_t> col(std::max(ptr[k], 1));
        for (int i = 0; i < ptr[k]; i++) col[i] = (idx_t)idx[i];
        CK(cudaMalloc(dptr, p.size() * sizeof(int)));
        CK(cudaMemcpy(*dptr, p.data(), p.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMalloc(dcol, col.size() * sizeof(idx_t)));
        CK(cudaMemcpy(*dcol, col.data(), col.size() * sizeof(idx_t), cudaMemcpyHostToDevice));
        return 0;
    };
    if (c->kx) if (int rc2 = up_rows(c->kx, lx_indptr, lx_indices, &c->lx_ptr, &c->lx_col)) return rc2;
    if (c->kz) if (int rc2 = up_rows(c->kz, lz_indptr, lz_indices, &c->lz_ptr, &c->lz_col)) return rc2;
    if (c->kx) {
        pack_rows(n, c->kx, lx_indptr, lx_indices, bits);
        CK(cudaMalloc(&c->lx_bits, bits.size() * 4));
        CK(cudaMemcpy(c->lx_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    }
    if (c->kz) {
        pack_rows(n, c->kz, lz_indptr, lz_indices, bits);
        CK(cudaMalloc(&c->lz_bits, bits.size() * 4));
        CK(cudaMemcpy(c->lz_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    }
    *out = c;
    return 0;
}

extern "C" int fbgnn_code_destroy(fbgnn_code *c) {
    if (!c) return 0;
    cudaSetDevice(c->ctx->device);
    fbgnn_graph_destroy(c->X);
    fbgnn_graph_destroy(c->Z);
    fbgnn_graph_destroy(c->basis_x);
    fbgnn_graph_destroy(c->basis_z);
    cudaFree(c->pivot_x);
    cudaFree(c->pivot_z);
    cudaFree(c->lx_ptr); cudaFree(c->lx_col); cudaFree(c->lz_ptr); cudaFree(c->lz_col);
    cudaFree(c->lx_bits);
    cudaFree(c->lz_bits);
    delete c;
    return 0;
}

static int make_basis(fbgnn_ctx *ctx, int n, const std::vector<int32_t> &ptr, const std::vector<int32_t> &idx,
                      int32_t rank, const int32_t *pivot, fbgnn_graph **graph, idx_t **dpivot) {
    const int m = (int)ptr.size() - 1;
    std::vector<int32_t> bp(1, 0), bi;
    std::vector<idx_t> piv(std::max(rank, 1));
    for (int r = 0; r < rank; r++) {
        REQUIRE(pivot[r] >= 0 && pivot[r] < m, "pivot row %d out of range", pivot[r]);
        bi.insert(bi.end(), idx.begin() + ptr[pivot[r]], idx.begin() + ptr[pivot[r] + 1]);
        bp.push_back((int32_t)bi.size());
        piv[r] = (idx_t)pivot[r];
    }
    if (int rc = fbgnn_graph_create(ctx, n, rank, bp.data(), bi.data(), graph)) return rc;
    CK(cudaMalloc(dpivot, piv.size() * sizeof(idx_t)));
    CK(cudaMemcpy(*dpivot, piv.data(), piv.size() * sizeof(idx_t), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int fbgnn_code_set_basis(fbgnn_code *code, int32_t rank_x, const int32_t *pivot_hx, int32_t rank_z,
                                    const int32_t *pivot_hz) {
    REQUIRE(code && pivot_hx && pivot_hz && rank_x > 0 && rank_z > 0, "bad argument");
    if (set_device(code->ctx)) return FBGNN_E_CUDA;
    fbgnn_graph_destroy(code->basis_x); code->basis_x = nullptr;
    fbgnn_graph_destroy(code->basis_z); code->basis_z = nullptr;
    cudaFree(code->pivot_x); code->pivot_x = nullptr;
    cudaFree(code->pivot_z); code->pivot_z = nullptr;
    const int n = code->X->dev.n;
    if (int rc = make_basis(code->ctx, n, code->hx_ptr, code->hx_idx, rank_x, pivot_hx, &code->basis_x, &code->pivot_x)) return rc;
    if (int rc = make_basis(code->ctx, n, code->hz_ptr, code->hz_idx, rank_z, pivot_hz, &code->basis_z, &code->pivot_z)) return rc;
    return 0;
}

extern "C" int fbgnn_code_edges(fbgnn_code *code, int32_t *e_x, int32_t *e_z) {
    REQUIRE(code, "code is NULL");
    if (e_x) *e_x = code->X->dev.E;
    if (e_z) *e_z = code->Z->dev.E;
    return 0;
}

// ------------------------------------------------------------------ launch helpers ------
// Threads per CTA for the one-frame-per-CTA kernels: the multiple of 32 in [128, 512] that
// wastes the fewest lanes over the variable-node and check-node passes.
static int pick_threads(int n_items_a, int n_items_b) {
    auto eff = [&](int t) {
        const double pa = (double)((n_items_a + t - 1) / t) * t, pb = (double)((n_items_b + t - 1) / t) * t;
        return (double)(n_items_a + n_items_b) / (pa + pb);
    };
    int best = 256;
    for (int t = 128; t <= 512; t += 32)
        if (eff(t) > eff(best) + 0.004) best = t;      // keep 256 unless another size is clearly better
    return best;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, fbgnn_ctx *ctx, const char *what) {
    if (bytes > ctx->smem_optin)
        return fail(FBGNN_E_UNSUPPORTED, "%s needs %zu bytes of shared memory per frame, the device offers %zu",
                    what, bytes, ctx->smem_optin);
    if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static size_t bp4_smem(const SideDev &X, const SideDev &Z, bool const_prior, bool iter_logits) {
    return sizeof(float) * ((size_t)X.E + Z.E + ((const_prior ? 2 : 3) + (iter_logits ? 2 : 0)) * (size_t)X.n) +
           2 * (((size_t)X.n + 1) & ~(size_t)1) + X.m + Z.m + X.n + 16;
}

template <bool CP, int DV, int DC, typename MATH, bool FPX>
static int launch_bp4_t(fbgnn_ctx *ctx, const Bp4Args &a, int64_t grid, size_t smem, int threads) {
    if (int rc = set_smem(k_bp4<CP, DV, DC, MATH, FPX>, smem, ctx, "quaternary BP")) return rc;
    k_bp4<CP, DV, DC, MATH, FPX><<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

template <int DV, int DC>
static int launch_bp4_m(fbgnn_ctx *ctx, const Bp4Args &a, int64_t grid, size_t smem, int threads, bool cp) {
    if (ctx->math_mode == FBGNN_MATH_FAST)
        return cp ? launch_bp4_t<true, DV, DC, MathFast, false>(ctx, a, grid, smem, threads)
                  : launch_bp4_t<false, DV, DC, MathFast, false>(ctx, a, grid, smem, threads);
    // fixed-point exit: exact arithmetic, regular graph, boxplus-phi, long runs (the bookkeeping costs ~5 %
    // per unsaturated iteration and a 16-iteration stage does not converge-and-saturate in time)
    const bool fpx = DV > 0 && a.cn_type == 0 && a.num_iter >= 32 && !a.iter_logits.ptr;
    if (fpx)
        return cp ? launch_bp4_t<true, DV, DC, MathExact, true>(ctx, a, grid, smem, threads)
                  : launch_bp4_t<false, DV, DC, MathExact, true>(ctx, a, grid, smem, threads);
    return cp ? launch_bp4_t<true, DV, DC, MathExact, false>(ctx, a, grid, smem, threads)
              : launch_bp4_t<false, DV, DC, MathExact, false>(ctx, a, grid, smem, threads);
}

// Codes whose per-frame state does not fit the shared memory of an SM: generic kernel with the float arrays in HBM.
template <typename MATH>
static int launch_bp4_gstate(fbgnn_ctx *ctx, Bp4Args a, int64_t grid, int threads, bool cp) {
    const SideDev &X = a.X, &Z = a.Z;
    const size_t smem = 2 * (((size_t)X.n + 1) & ~(size_t)1) + X.m + Z.m + X.n + 16;
    a.state_stride = (int64_t)X.E + Z.E + ((cp ? 2 : 3) + (a.iter_logits.ptr ? 2 : 0)) * (int64_t)X.n;
    CK(cudaMallocAsync(&a.state, sizeof(float) * (size_t)grid * a.state_stride, ctx->stream));
    int rc;
    if (cp) {
        rc = set_smem(k_bp4<true, 0, 0, MATH, false, true>, smem, ctx, "quaternary BP (HBM state)");
        if (!rc) k_bp4<true, 0, 0, MATH, false, true><<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    } else {
        rc = set_smem(k_bp4<false, 0, 0, MATH, false, true>, smem, ctx, "quaternary BP (HBM state)");
        if (!rc) k_bp4<false, 0, 0, MATH, false, true><<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    }
    CK(cudaFreeAsync(a.state, ctx->stream));
    if (rc) return rc;
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

static int launch_bp4(fbgnn_ctx *ctx, const Bp4Args &a, int64_t grid) {
    if (grid <= 0) return 0;
    const bool cp = a.llr.ptr == nullptr;
    const size_t smem = bp4_smem(a.X, a.Z, cp, a.iter_logits.ptr != nullptr);
    const int threads = pick_threads(a.X.n, a.X.m + a.Z.m);
    if (smem > ctx->smem_optin)
        return ctx->math_mode == FBGNN_MATH_FAST ? launch_bp4_gstate<MathFast>(ctx, a, grid, threads, cp)
                                                 : launch_bp4_gstate<MathExact>(ctx, a, grid, threads, cp);
    // both sides regular with the same degrees -> unrolled instantiation
    int dv = 0, dc = 0;
    if (a.X.reg_dv && a.X.reg_dv == a.Z.reg_dv && a.X.reg_dc && a.X.reg_dc == a.Z.reg_dc) { dv = a.X.reg_dv; dc = a.X.reg_dc; }
    if (dv == 3 && dc == 6) return launch_bp4_m<3, 6>(ctx, a, grid, smem, threads, cp);
    if (dv == 4 && dc == 8) return launch_bp4_m<4, 8>(ctx, a, grid, smem, threads, cp);
    if (dv == 5 && dc == 10) return launch_bp4_m<5, 10>(ctx, a, grid, smem, threads, cp);
    return launch_bp4_m<0, 0>(ctx, a, grid, smem, threads, cp);
}

// ------------------------------------------------------------------ noise sources -------
extern "C" int fbgnn_pauli_sample(fbgnn_ctx *ctx, int32_t n, int64_t B, const float thr[3], uint64_t seed,
                                  uint64_t first_frame, fbgnn_tensor2 noise_x, fbgnn_tensor2 noise_z) {
    REQUIRE(ctx && thr && n > 0 && B >= 0, "bad argument");
    REQUIRE(noise_x.ptr && noise_z.ptr, "noise outputs are NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (B == 0) return 0;
    SampleArgs a{};
    a.X.n = n; a.mode = 0;
    a.thr0 = thr[0]; a.thr1 = thr[1]; a.thr2 = thr[2];
    a.seed = seed; a.first_frame = first_frame;
    a.nx_out = v2<uint8_t>(noise_x); a.nz_out = v2<uint8_t>(noise_z);
    k_sample<<<(unsigned)B, 128, (size_t)n, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

extern "C" int fbgnn_pauli_sample_wt(fbgnn_ctx *ctx, int32_t n, int64_t B, int32_t wt, uint64_t seed,
                                     uint64_t first_frame, fbgnn_tensor2 noise_x, fbgnn_tensor2 noise_z) {
    REQUIRE(ctx && n > 0 && n <= 65535 && B >= 0 && wt >= 0, "bad argument");
    REQUIRE(noise_x.ptr && noise_z.ptr, "noise outputs are NULL");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (B == 0) return 0;
    SampleArgs a{};
    a.X.n = n; a.mode = 2; a.wt = wt;
    a.seed = seed; a.first_frame = first_frame;
    a.nx_out = v2<uint8_t>(noise_x); a.nz_out = v2<uint8_t>(noise_z);
    k_sample<<<(unsigned)B, 128, (size_t)3 * n + 8, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

extern "C" int fbgnn_bsc_sample(fbgnn_ctx *ctx, int32_t n, int64_t B, float p, uint64_t seed,
                                uint64_t first_frame, fbgnn_tensor2 noise) {
    REQUIRE(ctx && n > 0 && B >= 0 && noise.ptr, "bad argument");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (B == 0) return 0;
    SampleArgs a{};
    a.X.n = n; a.mode = 1; a.thr0 = p;
    a.seed = seed; a.first_frame = first_frame;
    a.nx_out = v2<uint8_t>(noise);
    k_sample<<<(unsigned)B, 128, (size_t)n, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

extern "C" int fbgnn_syndrome(fbgnn_graph *g, int64_t B, fbgnn_tensor2 noise, fbgnn_tensor2 syndrome) {
    REQUIRE(g && noise.ptr && syndrome.ptr && B >= 0, "bad argument");
    if (int rc = need_decodable(g)) return rc;
    fbgnn_ctx *ctx = g->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (B == 0) return 0;
    SyndromeArgs a{g->dev, v2<const uint8_t>(noise), v2<uint8_t>(syndrome)};
    k_syndrome<<<(unsigned)B, 128, (size_t)g->dev.n, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// ------------------------------------------------------------------ decoders ------------
extern "C" int fbgnn_bp4_decode(fbgnn_code *code, int32_t cn_type, int32_t num_iter, float factor, int64_t B,
                                fbgnn_tensor3 llr, float prior, fbgnn_tensor2 synd_x, fbgnn_tensor2 synd_z,
                                fbgnn_tensor2 Lx, fbgnn_tensor2 Ly, fbgnn_tensor2 Lz, fbgnn_tensor2 x_hat,
                                fbgnn_tensor2 z_hat, fbgnn_tensor2 x_logit, fbgnn_tensor2 z_logit,
                                fbgnn_tensor2 msg_x, fbgnn_tensor2 msg_z, fbgnn_tensor3 iter_logits) {
    REQUIRE(code, "code is NULL");
    REQUIRE(cn_type >= 0 && cn_type <= 2, "unknown cn_type %d", cn_type);
    REQUIRE(num_iter >= 0 && B >= 0, "num_iter and B must be non-negative");
    REQUIRE(synd_x.ptr && synd_z.ptr, "syndromes are NULL");
    fbgnn_ctx *ctx = code->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    Bp4Args a{};
    a.X = code->X->dev; a.Z = code->Z->dev;
    a.cn_type = cn_type; a.num_iter = num_iter; a.factor = factor;
    a.llr = v3<const float>(llr); a.prior = prior;
    a.sx = v2<const uint8_t>(synd_x); a.sz = v2<const uint8_t>(synd_z);
    a.Lx = v2<float>(Lx); a.Ly = v2<float>(Ly); a.Lz = v2<float>(Lz);
    a.xh = v2<uint8_t>(x_hat); a.zh = v2<uint8_t>(z_hat);
    a.xl = v2<float>(x_logit); a.zl = v2<float>(z_logit);
    a.msg_x = v2<float>(msg_x); a.msg_z = v2<float>(msg_z);
    a.iter_logits = v3<float>(iter_logits);
    return launch_bp4(ctx, a, B);
}

static size_t bp2_smem(const SideDev &S) { return sizeof(float) * ((size_t)S.E + S.n) + S.m + S.n + 16; }

template <int DV, int DC, typename MATH>
static int launch_bp2_t(fbgnn_ctx *ctx, const Bp2Args &a, int64_t B, size_t smem) {
    if (int rc = set_smem(k_bp2<DV, DC, MATH>, smem, ctx, "binary BP")) return rc;
    k_bp2<DV, DC, MATH><<<(unsigned)B, pick_threads(a.S.n, a.S.m), smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

template <int DV, int DC>
static int launch_bp2_m(fbgnn_ctx *ctx, const Bp2Args &a, int64_t B, size_t smem) {
    return ctx->math_mode == FBGNN_MATH_FAST ? launch_bp2_t<DV, DC, MathFast>(ctx, a, B, smem)
                                             : launch_bp2_t<DV, DC, MathExact>(ctx, a, B, smem);
}

static int launch_bp2(fbgnn_ctx *ctx, const Bp2Args &a, int64_t B) {
    if (B <= 0) return 0;
    const size_t smem = bp2_smem(a.S);
    if (a.S.reg_dv == 3 && a.S.reg_dc == 6) return launch_bp2_m<3, 6>(ctx, a, B, smem);
    if (a.S.reg_dv == 4 && a.S.reg_dc == 8) return launch_bp2_m<4, 8>(ctx, a, B, smem);
    if (a.S.reg_dv == 5 && a.S.reg_dc == 10) return launch_bp2_m<5, 10>(ctx, a, B, smem);
    return launch_bp2_m<0, 0>(ctx, a, B, smem);
}

extern "C" int fbgnn_bp2_decode(fbgnn_graph *g, int32_t cn_type, int32_t num_iter, float factor, int64_t B,
                                fbgnn_tensor2 llr, fbgnn_tensor2 synd, fbgnn_tensor2 soft, fbgnn_tensor2 hard) {
    REQUIRE(g, "graph is NULL");
    REQUIRE(cn_type >= 0 && cn_type <= 2, "unknown cn_type %d", cn_type);
    REQUIRE(num_iter >= 0 && B >= 0, "num_iter and B must be non-negative");
    REQUIRE(llr.ptr, "llr is NULL");
    REQUIRE(soft.ptr || hard.ptr, "no output requested");
    if (int rc = need_decodable(g)) return rc;
    fbgnn_ctx *ctx = g->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    Bp2Args a{};
    a.S = g->dev; a.cn_type = cn_type; a.num_iter = num_iter; a.factor = factor;
    a.llr = v2<const float>(llr); a.synd = v2<const uint8_t>(synd);
    a.soft = v2<float>(soft); a.hard = v2<uint8_t>(hard);
    return launch_bp2(ctx, a, B);
}

// ------------------------------------------------------------------ feedback GNN --------
template <int H, int M>
static void pack_gnn(std::vector<float> &w, const float *W0, const float *b0, const float *W1x,
                     const float *b1x, const float *W2x, const float *b2x, const float *W1z,
                     const float *b1z, const float *W2z, const float *b2z, const float *W3, const float *b3) {
    typedef GnnLayout<H, M> L;
    w.assign(L::total, 0.0f);
    auto put = [&](int off, const float *src, int count) { if (src) std::memcpy(&w[off], src, sizeof(float) * count); };
    put(L::W1x, W1x, 4 * H); put(L::b1x, b1x, H); put(L::W2x, W2x, H * M); put(L::b2x, b2x, M);
    put(L::W1z, W1z, 4 * H); put(L::b1z, b1z, H); put(L::W2z, W2z, H * M); put(L::b2z, b2z, M);
    put(L::W3, W3, (2 * M + 3) * H); put(L::b3, b3, H); put(L::W0, W0, H * 3); put(L::b0, b0, 3);
}

extern "C" int fbgnn_gnn_create(fbgnn_ctx *ctx, int32_t H, int32_t M, int32_t activation, int32_t reduce_op,
                                const float *W0, const float *b0, const float *W1x, const float *b1x,
                                const float *W2x, const float *b2x, const float *W1z, const float *b1z,
                                const float *W2z, const float *b2z, const float *W3, const float *b3,
                                fbgnn_gnn **out) {
    REQUIRE(ctx && out, "NULL argument");
    REQUIRE(W0 && W1x && W2x && W1z && W2z && W3, "weight matrices must not be NULL");
    REQUIRE(activation >= 0 && activation <= 2, "unknown activation %d", activation);
    REQUIRE(reduce_op >= 0 && reduce_op <= 3, "unknown reduce_op %d", reduce_op);
    const bool any_b = b0 || b1x || b2x || b1z || b2z || b3, all_b = b0 && b1x && b2x && b1z && b2z && b3;
    REQUIRE(any_b == all_b, "either all biases or none must be given");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    std::vector<float> w;
    if (H == 40 && M == 20) pack_gnn<40, 20>(w, W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3);
    else if (H == 20 && M == 20) pack_gnn<20, 20>(w, W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3);
    else if (H == 64 && M == 32) pack_gnn<64, 32>(w, W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3);
    else return fail(FBGNN_E_UNSUPPORTED, "Feedback_GNN with num_hidden_units=%d, num_msg_dims=%d is not "
                     "compiled into this build (available: 40/20, 20/20, 64/32)", H, M);
    fbgnn_gnn *g = new fbgnn_gnn();
    g->ctx = ctx; g->H = H; g->M = M; g->act = activation; g->reduce = reduce_op; g->use_bias = all_b ? 1 : 0;
    g->total = (int)w.size();
    CK(cudaMalloc(&g->weights, w.size() * sizeof(float)));
    CK(cudaMemcpy(g->weights, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    *out = g;
    return 0;
}

extern "C" int fbgnn_gnn_destroy(fbgnn_gnn *g) {
    if (!g) return 0;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->weights);
    delete g;
    return 0;
}

template <int H, int M, int DV, bool TB, bool FACT, typename MATH>
static int launch_gnn_t(fbgnn_ctx *ctx, const GnnArgs &a) {
    const size_t smem = sizeof(float) * GnnLayout<H, M>::total;
    if (int rc = set_smem(k_gnn<H, M, DV, TB, FACT, MATH>, smem, ctx, "feedback GNN")) return rc;
    const int64_t items = a.num_frames * a.X.n;
    int64_t blocks = (items + 127) / 128;
    blocks = std::min<int64_t>(blocks, (int64_t)ctx->num_sms * 8);
    k_gnn<H, M, DV, TB, FACT, MATH><<<(unsigned)blocks, 128, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

template <int H, int M, int DV, bool TB, bool FACT = true>
static int launch_gnn_m(fbgnn_ctx *ctx, const GnnArgs &a) {
    return ctx->math_mode == FBGNN_MATH_FAST ? launch_gnn_t<H, M, DV, TB, FACT, MathFast>(ctx, a)
                                             : launch_gnn_t<H, M, DV, TB, FACT, MathExact>(ctx, a);
}

static int launch_gnn(fbgnn_ctx *ctx, const fbgnn_gnn *g, GnnArgs &a) {
    if (a.num_frames <= 0) return 0;
    a.weights = g->weights; a.act = g->act; a.reduce = g->reduce; a.use_bias = g->use_bias;
    const bool reg3 = a.X.reg_dv == 3 && a.Z.reg_dv == 3;      // the (3,6)-regular GHP / bivariate codes
    const bool tb = g->act == FBGNN_ACT_TANH && g->use_bias;   // the shipped configuration
    const bool fact = g->reduce <= 1;                          // mean / sum: output layer after the reduction
    if (g->H == 40 && g->M == 20) {
        if (!fact) return launch_gnn_m<40, 20, 0, false, false>(ctx, a);
        if (reg3 && tb) return launch_gnn_m<40, 20, 3, true>(ctx, a);
        if (reg3) return launch_gnn_m<40, 20, 3, false>(ctx, a);
        return tb ? launch_gnn_m<40, 20, 0, true>(ctx, a) : launch_gnn_m<40, 20, 0, false>(ctx, a);
    }
    if (g->H == 20 && g->M == 20) return fact ? launch_gnn_m<20, 20, 0, false>(ctx, a) : launch_gnn_m<20, 20, 0, false, false>(ctx, a);
    if (g->H == 64 && g->M == 32) return fact ? launch_gnn_m<64, 32, 0, false>(ctx, a) : launch_gnn_m<64, 32, 0, false, false>(ctx, a);
    return fail(FBGNN_E_UNSUPPORTED, "unsupported GNN dimensions");
}

extern "C" int fbgnn_gnn_forward(fbgnn_code *code, fbgnn_gnn *gnn, int64_t B, fbgnn_tensor3 h_vn,
                                 fbgnn_tensor2 logit_hx, fbgnn_tensor2 logit_hz, fbgnn_tensor2 synd_x,
                                 fbgnn_tensor2 synd_z, fbgnn_tensor3 out) {
    REQUIRE(code && gnn, "NULL handle");
    REQUIRE(h_vn.ptr && logit_hx.ptr && logit_hz.ptr && synd_x.ptr && synd_z.ptr && out.ptr, "NULL tensor");
    REQUIRE(B >= 0, "B must be non-negative");
    fbgnn_ctx *ctx = code->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    GnnArgs a{};
    a.X = code->X->dev; a.Z = code->Z->dev;
    a.num_frames = B;
    a.h_vn = v3<const float>(h_vn);
    a.logit_hx = v2<const float>(logit_hx); a.logit_hz = v2<const float>(logit_hz);
    a.sx = v2<const uint8_t>(synd_x); a.sz = v2<const uint8_t>(synd_z);
    a.out = v3<float>(out);
    return launch_gnn(ctx, gnn, a);
}

// ------------------------------------------------------------------ OSD-0 ---------------
static int launch_osd0(fbgnn_ctx *ctx, Osd0Args &a, int64_t grid) {
    if (grid <= 0) return 0;
    const int n = a.S.n, R = a.S.m, W = (n + 1 + 31) / 32, Rp = R | 1;
    int npad = 1;
    while (npad < n) npad <<= 1;
    a.npad = npad;
    const size_t main_bytes = std::max<size_t>((size_t)npad * 8, (size_t)W * Rp * 4);
    const size_t smem = ((main_bytes + 7) & ~(size_t)7) + sizeof(uint16_t) * (2 * (size_t)n + R) + 16;
    if (int rc = set_smem(k_osd0, smem, ctx, "OSD-0")) return rc;
    k_osd0<<<(unsigned)grid, 256, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

extern "C" int fbgnn_osd0_decode(fbgnn_graph *basis, int64_t B, fbgnn_tensor2 llr, fbgnn_tensor2 synd,
                                 fbgnn_tensor2 e_hat) {
    REQUIRE(basis && llr.ptr && synd.ptr && e_hat.ptr && B >= 0, "bad argument");
    if (int rc = need_decodable(basis)) return rc;
    fbgnn_ctx *ctx = basis->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    Osd0Args a{};
    a.S = basis->dev;
    a.llr = v2<const float>(llr); a.sign = 1.0f;
    a.synd = v2<const uint8_t>(synd);
    a.e_hat = v2<uint8_t>(e_hat);
    return launch_osd0(ctx, a, B);
}

// ------------------------------------------------------------------ GNN_BP4 -------------
struct fbgnn_gbp {
    fbgnn_ctx *ctx;
    int d, H, M, act, reduce, use_bias;
    int gemm = FBGNN_GEMM_FMA;
    float *w_cn = nullptr, *w_vn = nullptr, *w_inv = nullptr;
    float *w_vn_tc = nullptr, *w_cn_tc[2] = {nullptr, nullptr};     // tensor-core operand tiles (fbgnn_gbp_tc.cuh)
};

typedef GbpLayout<20, 40, 20> GL;

static void pack_edge(std::vector<float> &w, int off, const float *W1, const float *b1, const float *W2, const float *b2) {
    const int D = 20, H = 40, M = 20;
    for (int j = 0; j < H; j++) for (int k = 0; k < 2 * D; k++) w[off + j * 2 * D + k] = W1[k * H + j];     // transposed
    if (b1) std::memcpy(&w[off + GL::e_b1], b1, sizeof(float) * H);
    std::memcpy(&w[off + GL::e_W2], W2, sizeof(float) * H * M);
    if (b2) std::memcpy(&w[off + GL::e_b2], b2, sizeof(float) * M);
}
static void pack_node(std::vector<float> &w, int off, int K, const float *W1, const float *b1, const float *W2, const float *b2) {
    const int D = 20, H = 40;
    std::memcpy(&w[off], W1, sizeof(float) * K * H);
    if (b1) std::memcpy(&w[off + GL::n_b1(K)], b1, sizeof(float) * H);
    std::memcpy(&w[off + GL::n_W2(K)], W2, sizeof(float) * H * D);
    if (b2) std::memcpy(&w[off + GL::n_b2(K)], b2, sizeof(float) * D);
}

extern "C" int fbgnn_gbp_destroy(fbgnn_gbp *g);

extern "C" int fbgnn_gbp_create(fbgnn_ctx *ctx, int32_t d, int32_t H, int32_t M, int32_t activation, int32_t reduce_op,
                                const float *const *arrays, fbgnn_gbp **out) {
    REQUIRE(ctx && arrays && out, "NULL argument");
    REQUIRE(activation >= 0 && activation <= 2 && reduce_op >= 0 && reduce_op <= 3, "bad activation / reduce_op");
    if (!(d == 20 && H == 40 && M == 20))
        return fail(FBGNN_E_UNSUPPORTED, "GNN_BP4 with embed/hidden/msg dims %d/%d/%d is not compiled into this build "
                    "(available: 20/40/20)", d, H, M);
    for (int i = 0; i < 30; i += 2) REQUIRE(arrays[i], "kernel %d is NULL", i / 2);
    bool any_b = false, all_b = true;
    for (int i = 1; i < 30; i += 2) { any_b |= arrays[i] != nullptr; all_b &= arrays[i] != nullptr; }
    REQUIRE(any_b == all_b, "either all biases or none must be given");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    // arrays: [Winv, binv, cn.msg_x (W1,b1,W2,b2), cn.msg_z, cn.embed_x, cn.embed_z, vn.msg_x, vn.msg_z, vn.embed]
    std::vector<float> wc(GL::cn_total, 0.0f), wv(GL::vn_total, 0.0f), wi(20 * 3 + 4, 0.0f);
    std::memcpy(wi.data(), arrays[0], sizeof(float) * 60);
    if (arrays[1]) std::memcpy(&wi[60], arrays[1], sizeof(float) * 3);
    const float *const *p = arrays + 2;
    pack_edge(wc, 0, p[0], p[1], p[2], p[3]);
    pack_edge(wc, GL::edge, p[4], p[5], p[6], p[7]);
    pack_node(wc, 2 * GL::edge, GL::KC, p[8], p[9], p[10], p[11]);
    pack_node(wc, 2 * GL::edge + GL::node(GL::KC), GL::KC, p[12], p[13], p[14], p[15]);
    pack_edge(wv, 0, p[16], p[17], p[18], p[19]);
    pack_edge(wv, GL::edge, p[20], p[21], p[22], p[23]);
    pack_node(wv, 2 * GL::edge, GL::KV, p[24], p[25], p[26], p[27]);
    fbgnn_gbp *g = new fbgnn_gbp();
    g->ctx = ctx; g->d = d; g->H = H; g->M = M; g->act = activation; g->reduce = reduce_op; g->use_bias = all_b ? 1 : 0;
    auto up = [&](const std::vector<float> &h, float **dptr) -> int {
        CK(cudaMalloc(dptr, h.size() * sizeof(float)));
        CK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    };
    if (up(wc, &g->w_cn) || up(wv, &g->w_vn) || up(wi, &g->w_inv)) { delete g; return FBGNN_E_CUDA; }
    {   // tensor-core operand tiles: TF32 hi / lo parts in the canonical K-major UMMA layout
        auto tile = [](std::vector<float> &buf, int off, int kpad, int npad, auto wf) {
            for (int nn = 0; nn < npad; nn++)
                for (int k = 0; k < kpad; k++) {
                    const float w = wf(k, nn), hi = tc::tf32_hi(w);
                    buf[off + tc::b_tile_offset(nn, k, kpad)] = hi;
                    buf[off + kpad * npad + tc::b_tile_offset(nn, k, kpad)] = w - hi;
                }
        };
        auto bias = [](std::vector<float> &buf, int off, const float *b, int cnt) {
            for (int i = 0; i < cnt; i++) buf[off + i] = b ? b[i] : 0.0f;
        };
        using tc::VnW; using tc::CnW;
        std::vector<float> tv(VnW::total, 0.0f);
        for (int sd = 0; sd < 2; sd++) {
            const float *W1 = p[16 + 4 * sd], *W2 = p[18 + 4 * sd];
            tile(tv, sd ? VnW::B1Z : VnW::B1X, 24, 48, [&](int k, int nn) { return (k < 20 && nn < 40) ? W1[(20 + k) * 40 + nn] : 0.0f; });
            tile(tv, sd ? VnW::W2Z : VnW::W2X, 40, 32, [&](int k, int nn) { return nn < 20 ? W2[k * 20 + nn] : 0.0f; });
            bias(tv, VnW::BIAS + sd * 40, p[17 + 4 * sd], 40);
            bias(tv, VnW::BIAS + 80 + sd * 20, p[19 + 4 * sd], 20);
        }
        tile(tv, VnW::W3AB, 40, 48, [&](int k, int nn) { return nn < 40 ? p[24][k * 40 + nn] : 0.0f; });
        tile(tv, VnW::W3C, 24, 48, [&](int k, int nn) { return (k < 20 && nn < 40) ? p[24][(40 + k) * 40 + nn] : 0.0f; });
        tile(tv, VnW::W4, 40, 32, [&](int k, int nn) { return nn < 20 ? p[26][k * 20 + nn] : 0.0f; });
        tile(tv, VnW::W5, 24, 80, [&](int k, int nn) { return k < 20 ? (nn < 40 ? p[0][k * 40 + nn] : p[4][k * 40 + nn - 40]) : 0.0f; });
        bias(tv, VnW::BIAS + 120, p[25], 40);
        bias(tv, VnW::BIAS + 160, p[27], 20);
        if (up(tv, &g->w_vn_tc)) { fbgnn_gbp_destroy(g); return FBGNN_E_CUDA; }
        for (int sd = 0; sd < 2; sd++) {
            std::vector<float> tcn(CnW::total, 0.0f);
            const float *mW1 = p[4 * sd], *mW2 = p[2 + 4 * sd], *eW1 = p[8 + 4 * sd], *eW2 = p[10 + 4 * sd], *vW1 = p[16 + 4 * sd];
            tile(tcn, CnW::B1, 24, 48, [&](int k, int nn) { return (k < 20 && nn < 40) ? mW1[(20 + k) * 40 + nn] : 0.0f; });
            tile(tcn, CnW::W2, 40, 32, [&](int k, int nn) { return nn < 20 ? mW2[k * 20 + nn] : 0.0f; });
            tile(tcn, CnW::W3A, 24, 48, [&](int k, int nn) { return nn >= 40 ? 0.0f : k < 20 ? eW1[k * 40 + nn] : k == 20 ? eW1[40 * 40 + nn] : 0.0f; });
            tile(tcn, CnW::W3B, 24, 48, [&](int k, int nn) { return (k < 20 && nn < 40) ? eW1[(20 + k) * 40 + nn] : 0.0f; });
            tile(tcn, CnW::W4, 40, 32, [&](int k, int nn) { return nn < 20 ? eW2[k * 20 + nn] : 0.0f; });
            tile(tcn, CnW::W5, 24, 48, [&](int k, int nn) { return (k < 20 && nn < 40) ? vW1[k * 40 + nn] : 0.0f; });
            bias(tcn, CnW::BIAS, p[1 + 4 * sd], 40);
            bias(tcn, CnW::BIAS + 40, p[3 + 4 * sd], 20);
            bias(tcn, CnW::BIAS + 60, p[9 + 4 * sd], 40);
            bias(tcn, CnW::BIAS + 100, p[11 + 4 * sd], 20);
            if (up(tcn, &g->w_cn_tc[sd])) { fbgnn_gbp_destroy(g); return FBGNN_E_CUDA; }
        }
    }
    *out = g;
    return 0;
}

extern "C" int fbgnn_gbp_destroy(fbgnn_gbp *g) {
    if (!g) return 0;
    cudaSetDevice(g->ctx->device);
    cudaFree(g->w_cn); cudaFree(g->w_vn); cudaFree(g->w_inv);
    cudaFree(g->w_vn_tc); cudaFree(g->w_cn_tc[0]); cudaFree(g->w_cn_tc[1]);
    delete g;
    return 0;
}

extern "C" int fbgnn_gbp_set_gemm(fbgnn_gbp *g, int32_t mode) {
    REQUIRE(g, "NULL handle");
    REQUIRE(mode == FBGNN_GEMM_FMA || mode == FBGNN_GEMM_TF32X3, "unknown gemm mode %d", mode);
    if (mode == FBGNN_GEMM_TF32X3 && !(g->reduce <= 1 && g->act == FBGNN_ACT_TANH))
        return fail(FBGNN_E_UNSUPPORTED, "the tensor-core path needs reduce_op mean / sum and tanh activation");
    g->gemm = mode;
    return 0;
}

template <typename MATH>
static int gbp_run(fbgnn_code *code, fbgnn_gbp *g, int32_t num_iter, int64_t B, GbpArgs a, fbgnn_tensor3 x_logit,
                   fbgnn_tensor3 z_logit, fbgnn_tensor2 x_hat, fbgnn_tensor2 z_hat) {
    fbgnn_ctx *ctx = code->ctx;
    cudaStream_t st = ctx->stream;
    const int n = a.X.n, mt = a.X.m + a.Z.m;
    const bool fact = g->reduce <= 1;        // mean / sum: factored edge MLPs (sender halves in a.pfc / a.pfv)
    const size_t smem_pre = sizeof(float) * 2 * 40 * 20;
    const size_t smem_codes = sizeof(uint32_t) * GBP_CODE_CAP * 128;          // staged edge codes of the factored kernels
    const size_t smem_cn = fact ? sizeof(float) * GL::cn_total + smem_pre + smem_codes : sizeof(float) * (GL::cn_total + 40 * 128);
    const size_t smem_vn = fact ? sizeof(float) * GL::vn_total + smem_pre + smem_codes : sizeof(float) * (GL::vn_total + 40 * 128);
    const bool tb = g->act == FBGNN_ACT_TANH && g->use_bias;
    if (fact) {
        if (int rc = set_smem(k_gbp_cn_f<20, 40, 20, true, MATH>, smem_cn, ctx, "GNN_BP4 CN update")) return rc;
        if (int rc = set_smem(k_gbp_vn_f<20, 40, 20, true, MATH>, smem_vn, ctx, "GNN_BP4 VN update")) return rc;
        if (int rc = set_smem(k_gbp_cn_f<20, 40, 20, false, MATH>, smem_cn, ctx, "GNN_BP4 CN update")) return rc;
        if (int rc = set_smem(k_gbp_vn_f<20, 40, 20, false, MATH>, smem_vn, ctx, "GNN_BP4 VN update")) return rc;
    } else {
        if (int rc = set_smem(k_gbp_cn<20, 40, 20, MATH>, smem_cn, ctx, "GNN_BP4 CN update")) return rc;
        if (int rc = set_smem(k_gbp_vn<20, 40, 20, MATH>, smem_vn, ctx, "GNN_BP4 VN update")) return rc;
    }
    const unsigned g_cn = (unsigned)std::min<int64_t>((B * mt + 127) / 128, (int64_t)ctx->num_sms * 8);
    const unsigned g_vn = (unsigned)std::min<int64_t>((B * n + 127) / 128, (int64_t)ctx->num_sms * 8);
    const size_t smem_lg = sizeof(float) * 2 * n + n + 16;
    const bool use_tc = fact && g->gemm == FBGNN_GEMM_TF32X3 && g->act == FBGNN_ACT_TANH;
    const size_t smem_vn_tc = sizeof(float) * tc::VnW::total + sizeof(uint32_t) * GBP_CODE_CAP * 256;
    const size_t smem_cn_tc = sizeof(float) * tc::CnW::total + sizeof(uint32_t) * GBP_CODE_CAP * 256;
    if (use_tc) {
        if (int rc = set_smem(tc::k_gbp_vn_tc<MATH>, smem_vn_tc, ctx, "GNN_BP4 VN update (tensor cores)")) return rc;
        if (int rc = set_smem(tc::k_gbp_cn_tc<MATH>, smem_cn_tc, ctx, "GNN_BP4 CN update (tensor cores)")) return rc;
    }
    auto tc_grid = [&](int64_t rows) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, (int64_t)ctx->num_sms * 2)); };
    auto cn_update = [&]() {
        if (use_tc) {
            for (int sd = 0; sd < 2; sd++) {
                const int ms = sd ? a.Z.m : a.X.m;
                if (ms == 0) continue;
                tc::k_gbp_cn_tc<MATH><<<tc_grid(B * ms), 256, smem_cn_tc, st>>>(a, g->w_cn_tc[sd], sd);
                ctx->launches++;
            }
            return;
        }
        if (fact && tb) k_gbp_cn_f<20, 40, 20, true, MATH><<<g_cn, 128, smem_cn, st>>>(a);
        else if (fact) k_gbp_cn_f<20, 40, 20, false, MATH><<<g_cn, 128, smem_cn, st>>>(a);
        else k_gbp_cn<20, 40, 20, MATH><<<g_cn, 128, smem_cn, st>>>(a);
        ctx->launches++;
    };
    if (fact) {
        k_gbp_pre_vn<20, 40, 20><<<g_vn, 128, smem_pre, st>>>(a);
        ctx->launches++;
    }
    a.zero_logits = 1;
    cn_update();
    a.zero_logits = 0;
    for (int it = 0; it < num_iter; it++) {
        if (use_tc) tc::k_gbp_vn_tc<MATH><<<tc_grid(B * n), 256, smem_vn_tc, st>>>(a, g->w_vn_tc);
        else if (fact && tb) k_gbp_vn_f<20, 40, 20, true, MATH><<<g_vn, 128, smem_vn, st>>>(a);
        else if (fact) k_gbp_vn_f<20, 40, 20, false, MATH><<<g_vn, 128, smem_vn, st>>>(a);
        else k_gbp_vn<20, 40, 20, MATH><<<g_vn, 128, smem_vn, st>>>(a);
        GbpArgs la = a;
        if (x_logit.ptr) la.x_logit = View2<float>{(float *)x_logit.ptr + it * x_logit.s0, x_logit.s1, x_logit.s2};
        if (z_logit.ptr) la.z_logit = View2<float>{(float *)z_logit.ptr + it * z_logit.s0, z_logit.s1, z_logit.s2};
        if (it == num_iter - 1) { la.x_hat = v2<uint8_t>(x_hat); la.z_hat = v2<uint8_t>(z_hat); }
        k_gbp_logit<20, MATH><<<(unsigned)B, 256, smem_lg, st>>>(la);
        ctx->launches += 2;
        if (it == num_iter - 1) break;
        cn_update();
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int fbgnn_gbp_decode(fbgnn_code *code, fbgnn_gbp *g, int32_t num_iter, int64_t B, fbgnn_tensor2 synd_x,
                                fbgnn_tensor2 synd_z, fbgnn_tensor3 x_logit, fbgnn_tensor3 z_logit,
                                fbgnn_tensor2 x_hat, fbgnn_tensor2 z_hat) {
    REQUIRE(code && g && synd_x.ptr && synd_z.ptr && x_hat.ptr && z_hat.ptr, "NULL argument");
    REQUIRE(num_iter >= 1 && B >= 0, "num_iter must be >= 1");
    REQUIRE(synd_x.s1 == 1 && synd_z.s1 == 1 && synd_x.s0 == code->X->dev.m && synd_z.s0 == code->Z->dev.m,
            "syndromes must be contiguous [B, m] (batch first)");
    fbgnn_ctx *ctx = code->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (B == 0) return 0;
    const SideDev &X = code->X->dev, &Z = code->Z->dev;
    const int n = X.n, D = 20, H = 40, mt = std::max(X.m + Z.m, 1);
    const bool fact = g->reduce <= 1;
    // The embeddings (and, factored, the sender halves) of a frame take (n + m) (D + H) floats + n H floats of
    // HBM: the batch is walked in chunks that keep this state to a few GB.
    const int64_t per_frame = sizeof(float) * ((int64_t)(n + mt) * D + (fact ? (int64_t)(mt + 2 * n) * H : 0) + mt);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)6 << 30) / per_frame));
    float *h_vn = nullptr, *hcx = nullptr, *hcz = nullptr, *lg = nullptr, *pfc = nullptr, *pfv = nullptr;
    cudaStream_t st = ctx->stream;
    CK(cudaMallocAsync(&h_vn, sizeof(float) * chunk * n * D, st));
    CK(cudaMallocAsync(&hcx, sizeof(float) * chunk * std::max(X.m, 1) * D, st));
    CK(cudaMallocAsync(&hcz, sizeof(float) * chunk * std::max(Z.m, 1) * D, st));
    CK(cudaMallocAsync(&lg, sizeof(float) * chunk * mt, st));
    if (fact) {
        CK(cudaMallocAsync(&pfc, sizeof(float) * chunk * mt * H, st));
        CK(cudaMallocAsync(&pfv, sizeof(float) * chunk * n * 2 * H, st));
    }
    int rc = 0;
    for (int64_t b0 = 0; b0 < B && rc == 0; b0 += chunk) {
        const int64_t nb = std::min(chunk, B - b0);
        k_fill<<<ctx->num_sms * 4, 256, 0, st>>>((uint32_t *)h_vn, nb * n * D, 0x3f800000u);   // h_vn = 1 (gnn.py:394)
        CK(cudaMemsetAsync(hcx, 0, sizeof(float) * nb * X.m * D, st));                          // h_cn = 0 (392-393)
        CK(cudaMemsetAsync(hcz, 0, sizeof(float) * nb * Z.m * D, st));
        ctx->launches++;
        GbpArgs a{};
        a.X = X; a.Z = Z; a.w_cn = g->w_cn; a.w_vn = g->w_vn; a.w_inv = g->w_inv;
        a.act = g->act; a.reduce = g->reduce; a.use_bias = g->use_bias; a.B = nb;
        a.h_vn = h_vn; a.hcx = hcx; a.hcz = hcz; a.lg = lg; a.pfc = pfc; a.pfv = pfv;
        a.sx = (const uint8_t *)synd_x.ptr + b0 * synd_x.s0; a.sz = (const uint8_t *)synd_z.ptr + b0 * synd_z.s0;
        a.lx_ptr = code->lx_ptr; a.lz_ptr = code->lz_ptr; a.lx_col = code->lx_col; a.lz_col = code->lz_col;
        a.kx = code->kx; a.kz = code->kz;
        fbgnn_tensor3 xl = x_logit, zl = z_logit;               // (iteration, row, frame)
        fbgnn_tensor2 xh = x_hat, zh = z_hat;                   // (qubit, frame)
        if (xl.ptr) xl.ptr = (float *)xl.ptr + b0 * xl.s2;
        if (zl.ptr) zl.ptr = (float *)zl.ptr + b0 * zl.s2;
        xh.ptr = (uint8_t *)xh.ptr + b0 * xh.s1;
        zh.ptr = (uint8_t *)zh.ptr + b0 * zh.s1;
        rc = ctx->math_mode == FBGNN_MATH_FAST ? gbp_run<MathFast>(code, g, num_iter, nb, a, xl, zl, xh, zh)
                                               : gbp_run<MathExact>(code, g, num_iter, nb, a, xl, zl, xh, zh);
    }
    CK(cudaFreeAsync(h_vn, st)); CK(cudaFreeAsync(hcx, st)); CK(cudaFreeAsync(hcz, st)); CK(cudaFreeAsync(lg, st));
    if (fact) { CK(cudaFreeAsync(pfc, st)); CK(cudaFreeAsync(pfv, st)); }
    return rc;
}

// ------------------------------------------------------------------ training: second stage
static int atb(fbgnn_ctx *ctx, const float *A, const float *Bm, int64_t R, int Ka, int Kb, float *partial, float *C) {
    if (Ka * Kb > 2048) return fail(FBGNN_E_UNSUPPORTED, "A^T B output too large");
    const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>((R + 255) / 256, 2 * (int64_t)ctx->num_sms));
    const size_t smem = sizeof(float) * 32 * (Ka + Kb);
    train::k_atb_partial<<<nblocks, 256, smem, ctx->stream>>>(A, Bm, R, Ka, Kb, partial);
    train::k_atb_reduce<<<(Ka * Kb + 255) / 256, 256, 0, ctx->stream>>>(partial, nblocks, Ka * Kb, C);
    CK(cudaGetLastError());
    ctx->launches += 2;
    return 0;
}

extern "C" int fbgnn_second_stage_grad(fbgnn_code *code, fbgnn_gnn *gnn, int32_t num_iter, float factor, int32_t loss_from,
                                       int64_t B, fbgnn_tensor3 h_vn, fbgnn_tensor2 logit_hx, fbgnn_tensor2 logit_hz,
                                       fbgnn_tensor2 synd_x, fbgnn_tensor2 synd_z, int32_t want_grad, double *loss,
                                       float *grads) {
    REQUIRE(code && gnn && loss, "NULL argument");
    REQUIRE(h_vn.ptr && logit_hx.ptr && logit_hz.ptr && synd_x.ptr && synd_z.ptr, "NULL tensor");
    REQUIRE(B > 0 && num_iter >= 1 && loss_from >= 0 && loss_from <= num_iter, "bad B / num_iter / loss_from");
    REQUIRE(!want_grad || grads, "grads buffer missing");
    if (!(gnn->H == 40 && gnn->M == 20 && gnn->act == FBGNN_ACT_TANH && gnn->use_bias && gnn->reduce <= 1))
        return fail(FBGNN_E_UNSUPPORTED, "gradients are provided for the 40 / 20 tanh GNN with bias and reduce mean / sum");
    fbgnn_ctx *ctx = code->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const SideDev &X = code->X->dev, &Z = code->Z->dev;
    const int n = X.n, E = X.E + Z.E, mt = X.m + Z.m, H = 40, M = 20;
    const int64_t R = B * n;
    cudaStream_t st = ctx->stream;
    const size_t smem_bp = sizeof(float) * ((size_t)3 * E + 14 * (size_t)n + mt);
    if (int rc = set_smem(train::k_bp4_grad, smem_bp, ctx, "BP4 gradient")) return rc;

    // one arena: new priors, BP trace, d priors, per-frame loss, row factors, partials, gradients
    const size_t n_llr = (size_t)R * 3, n_trace = (size_t)B * num_iter * E, n_dpr = (size_t)R * 3;
    const size_t n_rows = want_grad ? (size_t)R * (41 + 40 + 44 + 3 + 2 * 41 + 2 * 20) + (size_t)B * E * 45 : 0;
    const int nblk = 2 * ctx->num_sms;
    const size_t n_part = want_grad ? (size_t)nblk * 2048 : 0, n_grad = 41 * 3 + 2 * (5 * 40 + 41 * 20) + 44 * 40;
    float *arena = nullptr;
    double *dloss = nullptr;
    CK(cudaMallocAsync(&arena, sizeof(float) * (n_llr + n_trace + n_dpr + n_rows + n_part + n_grad), st));
    CK(cudaMallocAsync(&dloss, sizeof(double) * B, st));
    float *llr = arena, *trace = llr + n_llr, *dpr = trace + n_trace, *rows = dpr + n_dpr, *part = rows + n_rows, *gout = part + n_part;

    GnnArgs ga{};
    ga.X = X; ga.Z = Z; ga.num_frames = B;
    ga.h_vn = v3<const float>(h_vn);
    ga.logit_hx = v2<const float>(logit_hx); ga.logit_hz = v2<const float>(logit_hz);
    ga.sx = v2<const uint8_t>(synd_x); ga.sz = v2<const uint8_t>(synd_z);
    ga.out = View3<float>{llr, (int64_t)3 * n, 3, 1};                     // (b, v, k)
    int rc = launch_gnn(ctx, gnn, ga);

    train::Bp4GradArgs ba{};
    ba.X = X; ba.Z = Z; ba.num_iter = num_iter; ba.loss_from = loss_from; ba.factor = factor; ba.B = B;
    ba.llr = View3<const float>{llr, (int64_t)3 * n, 1, 3};               // (b, k, v) view of the GNN output
    ba.sx = ga.sx; ba.sz = ga.sz;
    ba.trace = trace; ba.dprior = want_grad ? dpr : nullptr; ba.loss = dloss;
    ba.wx = 1.0f / ((float)B * (float)std::max(Z.m, 1)); ba.wz = 1.0f / ((float)B * (float)std::max(X.m, 1));
    if (!rc) {
        train::k_bp4_grad<<<(unsigned)B, 256, smem_bp, st>>>(ba);
        ctx->launches++;
    }
    if (!rc && want_grad) {
        train::GnnBwdArgs wa{};
        wa.X = X; wa.Z = Z; wa.weights = gnn->weights; wa.reduce = gnn->reduce; wa.B = B;
        wa.h_vn = ga.h_vn; wa.logit_hx = ga.logit_hx; wa.logit_hz = ga.logit_hz; wa.sx = ga.sx; wa.sz = ga.sz;
        wa.dout = dpr;
        float *q = rows;
        wa.hid1 = q; q += (size_t)R * 41;
        wa.dpre3 = q; q += (size_t)R * 40;
        wa.in1 = q; q += (size_t)R * 44;
        wa.dout_r = q; q += (size_t)R * 3;
        wa.hs1 = q; q += (size_t)R * 2 * 41;
        wa.dm = q; q += (size_t)R * 2 * 20;
        wa.ftx = q; q += (size_t)B * X.E * 5;
        wa.dpx = q; q += (size_t)B * X.E * 40;
        wa.ftz = q; q += (size_t)B * Z.E * 5;
        wa.dpz = q;
        const size_t smem_w = sizeof(float) * GnnLayout<40, 20>::total;
        rc = set_smem(train::k_gnn_bwd_rows<40, 20>, smem_w, ctx, "feedback GNN gradient");
        if (!rc) {
            const unsigned blocks = (unsigned)std::min<int64_t>((R + 127) / 128, (int64_t)ctx->num_sms * 8);
            train::k_gnn_bwd_rows<40, 20><<<blocks, 128, smem_w, st>>>(wa);
            ctx->launches++;
        }
        // [W0; b0] | [W1x; b1x] | [W2x; b2x] | [W1z; b1z] | [W2z; b2z] | [W3; b3]
        float *g = gout;
        if (!rc) rc = atb(ctx, wa.hid1, wa.dout_r, R, 41, 3, part, g);
        g += 41 * 3;
        if (!rc) rc = atb(ctx, wa.ftx, wa.dpx, B * X.E, 5, 40, part, g);
        g += 5 * 40;
        if (!rc) rc = atb(ctx, wa.hs1, wa.dm, R, 41, 20, part, g);
        g += 41 * 20;
        if (!rc) rc = atb(ctx, wa.ftz, wa.dpz, B * Z.E, 5, 40, part, g);
        g += 5 * 40;
        if (!rc) rc = atb(ctx, wa.hs1 + (size_t)R * 41, wa.dm + (size_t)R * 20, R, 41, 20, part, g);
        g += 41 * 20;
        if (!rc) rc = atb(ctx, wa.in1, wa.dpre3, R, 44, 40, part, g);
        if (!rc) CK(cudaMemcpyAsync(grads, gout, sizeof(float) * n_grad, cudaMemcpyDeviceToHost, st));
    }
    std::vector<double> hl((size_t)B);
    if (!rc) CK(cudaMemcpyAsync(hl.data(), dloss, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFreeAsync(arena, st)); CK(cudaFreeAsync(dloss, st));
    if (rc) return rc;
    CK(cudaGetLastError());
    double sum = 0.0;
    for (double x : hl) sum += x;
    *loss = sum;
    return 0;
}

// ------------------------------------------------------------------ pipelines -----------
static int ws_reserve(fbgnn_ctx *ctx, int64_t B, int n, int m) {
    Workspace &w = ctx->ws;
    if (w.cap_frames >= B && w.n == n && w.m == m) return 0;
    CK(cudaStreamSynchronize(ctx->stream));
    ws_free(w);
    const size_t b = (size_t)B;
    CK(cudaMalloc(&w.vbits, b * n));
    CK(cudaMalloc(&w.sbits, b * std::max(m, 1)));
    CK(cudaMalloc(&w.active[0], b));
    CK(cudaMalloc(&w.active[1], b));
    CK(cudaMalloc(&w.rounds, b));
    CK(cudaMalloc(&w.L, b * 3 * n * sizeof(float)));
    CK(cudaMalloc(&w.P, b * 3 * n * sizeof(float)));
    CK(cudaMalloc(&w.logit, b * std::max(m, 1) * sizeof(float)));
    CK(cudaMalloc(&w.list[0], b * sizeof(int)));
    CK(cudaMalloc(&w.list[1], b * sizeof(int)));
    CK(cudaMalloc(&w.list_count, 2 * sizeof(int)));
    CK(cudaMalloc(&w.counters, 4 * sizeof(unsigned long long)));
    w.cap_frames = B; w.n = n; w.m = m;
    return 0;
}

extern "C" int fbgnn_pipeline_run(fbgnn_code *code, const fbgnn_pipeline_cfg *cfg, uint64_t seed,
                                  uint64_t first_frame, int64_t B, fbgnn_tensor2 noise_x, fbgnn_tensor2 noise_z,
                                  uint8_t *flags, fbgnn_tensor2 x_diff, fbgnn_tensor2 z_diff, int64_t *counters) {
    REQUIRE(code && cfg, "NULL argument");
    REQUIRE(cfg->num_stages >= 1 && cfg->num_iter && cfg->factor && cfg->cn_type, "bad pipeline configuration");
    REQUIRE(cfg->num_stages == 1 || cfg->gnn, "feedback GNNs missing");
    REQUIRE((noise_x.ptr == nullptr) == (noise_z.ptr == nullptr), "give both noise_x and noise_z or neither");
    REQUIRE(B >= 0 && B < ((int64_t)1 << 31), "bad batch size");
    REQUIRE(!cfg->osd0 || (code->basis_x && code->basis_z), "OSD-0 needs fbgnn_code_set_basis first");
    for (int s = 0; s < cfg->num_stages; s++) {
        REQUIRE(cfg->cn_type[s] >= 0 && cfg->cn_type[s] <= 2, "unknown cn_type in stage %d", s);
        REQUIRE(cfg->num_iter[s] >= 0, "negative num_iter in stage %d", s);
        REQUIRE(s == 0 || cfg->gnn[s - 1], "feedback GNN %d is NULL", s - 1);
    }
    fbgnn_ctx *ctx = code->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const SideDev &X = code->X->dev, &Z = code->Z->dev;
    const int n = X.n, m = X.m + Z.m;
    if (counters) std::memset(counters, 0, 4 * sizeof(int64_t));
    if (B == 0) return 0;
    if (int rc = ws_reserve(ctx, B, n, m)) return rc;
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(w.rounds, 0, (size_t)B, st));
    CK(cudaMemsetAsync(w.counters, 0, 4 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(w.list_count, 0, 2 * sizeof(int), st));

    // noise + syndromes
    SampleArgs sa{};
    sa.X = X; sa.Z = Z; sa.mode = cfg->fixed_weight > 0 ? 2 : 0; sa.wt = cfg->fixed_weight;
    sa.thr0 = cfg->thr[0]; sa.thr1 = cfg->thr[1]; sa.thr2 = cfg->thr[2];
    sa.seed = seed; sa.first_frame = first_frame;
    sa.nx_in = v2<const uint8_t>(noise_x); sa.nz_in = v2<const uint8_t>(noise_z);
    sa.vbits = w.vbits; sa.sbits = w.sbits;
    k_sample<<<(unsigned)B, 128, (size_t)3 * n + 8, st>>>(sa);
    CK(cudaGetLastError());
    ctx->launches++;

    const int S = cfg->num_stages;
    int64_t cur_count = B;            // frames the current stage runs on
    const int *cur_list = nullptr;
    for (int s = 0; s < S; s++) {
        const bool last = (s == S - 1);
        if (s > 0) {
            // feedbacks[s-1]: priors P from the marginals L and the soft syndromes
            GnnArgs ga{};
            ga.X = X; ga.Z = Z;
            ga.frame_list = cur_list; ga.num_frames = cur_count;
            ga.h_vn = View3<const float>{w.L, 3 * (int64_t)n, 1, n};
            ga.logit_hx = View2<const float>{w.logit, 1, m};           // z_logit: rows of hx
            ga.logit_hz = View2<const float>{w.logit + X.m, 1, m};     // x_logit: rows of hz
            ga.sx = View2<const uint8_t>{w.sbits, 1, m};
            ga.sz = View2<const uint8_t>{w.sbits + X.m, 1, m};
            ga.out = View3<float>{w.P, 3 * (int64_t)n, 1, n};
            if (int rc = launch_gnn(ctx, cfg->gnn[s - 1], ga)) return rc;
        }
        Bp4Args a{};
        a.X = X; a.Z = Z;
        a.cn_type = cfg->cn_type[s]; a.num_iter = cfg->num_iter[s]; a.factor = cfg->factor[s];
        a.frame_list = cur_list;
        if (s > 0) a.llr = View3<const float>{w.P, 3 * (int64_t)n, n, 1};
        a.prior = cfg->prior;
        a.sx = View2<const uint8_t>{w.sbits, 1, m};
        a.sz = View2<const uint8_t>{w.sbits + X.m, 1, m};
        if (!last || cfg->osd0) {
            a.Lx = View2<float>{w.L, 3 * (int64_t)n, 1};
            a.Ly = View2<float>{w.L + n, 3 * (int64_t)n, 1};
            a.Lz = View2<float>{w.L + 2 * n, 3 * (int64_t)n, 1};
            if (!last) {
                a.zl = View2<float>{w.logit, 1, m};
                a.xl = View2<float>{w.logit + X.m, 1, m};
            }
        }
        a.vbits = w.vbits;
        a.active_in = (s == 0) ? nullptr : w.active[(s - 1) & 1];
        a.active_out = w.active[s & 1];
        a.rounds = last ? nullptr : w.rounds;
        const bool compact = (cfg->skip_inactive && !last) || (last && cfg->osd0);
        if (compact) {
            CK(cudaMemsetAsync(w.list_count + (s & 1), 0, sizeof(int), st));
            a.next_list = w.list[s & 1];
            a.next_count = w.list_count + (s & 1);
        }
        if (int rc = launch_bp4(ctx, a, cur_count)) return rc;
        if (compact) {
            int cnt = 0;
            CK(cudaMemcpyAsync(&cnt, w.list_count + (s & 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            cur_count = cnt;
            cur_list = w.list[s & 1];
            if (cur_count == 0) break;
        }
    }

    if (cfg->osd0 && cur_count > 0 && cur_list) {
        // BP4_OSD_Model (bp_osd.py:80-197): frames still mismatching get both parts re-solved by OSD-0
        OsdLlrArgs la{n, cur_list, cur_count, w.L, w.P};
        const int64_t blocks = std::min<int64_t>((cur_count * n + 255) / 256, (int64_t)ctx->num_sms * 8);
        if (ctx->math_mode == FBGNN_MATH_FAST) k_osd_llr<MathFast><<<(unsigned)blocks, 256, 0, st>>>(la);
        else k_osd_llr<MathExact><<<(unsigned)blocks, 256, 0, st>>>(la);
        CK(cudaGetLastError());
        ctx->launches++;
        Osd0Args oa{};
        oa.frame_list = cur_list; oa.sign = 1.0f;
        oa.vbits = w.vbits;
        oa.S = code->basis_x->dev;                                   // hx basis, osd_llrz, syndrome_x -> z_hat
        oa.llr = View2<const float>{w.P + n, 3 * (int64_t)n, 1};
        oa.synd = View2<const uint8_t>{w.sbits, 1, m}; oa.synd_row = code->pivot_x; oa.vbit = 3;
        if (int rc = launch_osd0(ctx, oa, cur_count)) return rc;
        oa.S = code->basis_z->dev;                                   // hz basis, osd_llrx, syndrome_z -> x_hat
        oa.llr = View2<const float>{w.P, 3 * (int64_t)n, 1};
        oa.synd = View2<const uint8_t>{w.sbits + X.m, 1, m}; oa.synd_row = code->pivot_z; oa.vbit = 2;
        if (int rc = launch_osd0(ctx, oa, cur_count)) return rc;
    }

    FinalArgs fa{};
    fa.X = X; fa.Z = Z;
    fa.lx_bits = code->lx_bits; fa.lz_bits = code->lz_bits; fa.kx = code->kx; fa.kz = code->kz;
    fa.binary = 0;
    fa.vbits = w.vbits; fa.rounds = w.rounds; fa.flags = flags;
    fa.x_diff = v2<uint8_t>(x_diff); fa.z_diff = v2<uint8_t>(z_diff);
    fa.counters = w.counters;
    const int W = (n + 31) / 32;
    k_final<<<(unsigned)B, 128, (size_t)((n + 3) & ~3) + 8 * (size_t)W, st>>>(fa);
    CK(cudaGetLastError());
    ctx->launches++;
    if (counters) {
        unsigned long long h[4];
        CK(cudaMemcpyAsync(h, w.counters, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 4; i++) counters[i] = (int64_t)h[i];
    }
    return 0;
}

extern "C" int fbgnn_bsc_pipeline_run(fbgnn_graph *g, fbgnn_graph *logical, int32_t cn_type, int32_t num_iter,
                                      float factor, float llr_const, float p, uint64_t seed, uint64_t first_frame,
                                      int64_t B, fbgnn_tensor2 noise, uint8_t *flags, int64_t *counters,
                                      fbgnn_graph *osd_basis, const int32_t *osd_pivot) {
    REQUIRE(g, "graph is NULL");
    REQUIRE(!osd_basis || (osd_pivot && osd_basis->dev.n == g->dev.n && osd_basis->dev.m <= g->dev.m),
            "bad OSD-0 basis");
    REQUIRE(cn_type >= 0 && cn_type <= 2, "unknown cn_type %d", cn_type);
    REQUIRE(num_iter >= 0 && B >= 0 && B < ((int64_t)1 << 31), "bad argument");
    REQUIRE(!logical || logical->dev.n == g->dev.n, "logical_pcm has %d columns, pcm has %d",
            logical ? logical->dev.n : 0, g->dev.n);
    if (int rc = need_decodable(g)) return rc;
    fbgnn_ctx *ctx = g->ctx;
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const SideDev &S = g->dev;
    const int n = S.n, m = S.m;
    if (counters) std::memset(counters, 0, 4 * sizeof(int64_t));
    if (B == 0) return 0;
    if (int rc = ws_reserve(ctx, B, n, m)) return rc;
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(w.counters, 0, 4 * sizeof(unsigned long long), st));
    SampleArgs sa{};
    sa.X = S; sa.mode = 1; sa.thr0 = p;
    sa.seed = seed; sa.first_frame = first_frame;
    sa.nx_in = v2<const uint8_t>(noise);
    sa.vbits = w.vbits; sa.sbits = w.sbits;
    k_sample<<<(unsigned)B, 128, (size_t)n, st>>>(sa);
    CK(cudaGetLastError());
    ctx->launches++;
    Bp2Args a{};
    a.S = S; a.cn_type = cn_type; a.num_iter = num_iter; a.factor = factor;
    a.llr_const = llr_const;
    a.synd = View2<const uint8_t>{w.sbits, 1, m};
    a.vbits = w.vbits;                 // decision -> bit 2
    if (osd_basis) {
        CK(cudaMemsetAsync(w.list_count, 0, sizeof(int), st));
        a.soft = View2<float>{w.L, n, 1};
        a.next_list = w.list[0]; a.next_count = w.list_count;
    }
    if (int rc = launch_bp2(ctx, a, B)) return rc;
    if (osd_basis) {
        // BP2_OSD_Model (bp_osd.py:199-274): OSD-0 on the frames whose decision misses the syndrome
        if (int rc = need_decodable(osd_basis)) return rc;
        int cnt = 0;
        CK(cudaMemcpyAsync(&cnt, w.list_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (cnt > 0) {
            const int R = osd_basis->dev.m;
            std::vector<idx_t> piv(R);
            for (int r = 0; r < R; r++) {
                REQUIRE(osd_pivot[r] >= 0 && osd_pivot[r] < m, "pivot row out of range");
                piv[r] = (idx_t)osd_pivot[r];
            }
            idx_t *dp = nullptr;
            CK(cudaMallocAsync(&dp, R * sizeof(idx_t), st));
            CK(cudaMemcpyAsync(dp, piv.data(), R * sizeof(idx_t), cudaMemcpyHostToDevice, st));
            Osd0Args oa{};
            oa.S = osd_basis->dev; oa.frame_list = w.list[0];
            oa.llr = View2<const float>{w.L, n, 1}; oa.sign = -1.0f;      // llr_hat = -decoder output
            oa.synd = View2<const uint8_t>{w.sbits, 1, m}; oa.synd_row = dp;
            oa.vbits = w.vbits; oa.vbit = 2;
            if (int rc = launch_osd0(ctx, oa, cnt)) return rc;
            CK(cudaStreamSynchronize(st));                                  // piv must outlive the copy
            CK(cudaFreeAsync(dp, st));
        }
    }
    FinalArgs fa{};
    fa.X = S;
    fa.lx_bits = logical ? logical->dev.bitrows : nullptr;
    fa.kx = logical ? logical->dev.m : 0;
    fa.binary = 1;
    fa.vbits = w.vbits; fa.flags = flags; fa.counters = w.counters;
    const int W = (n + 31) / 32;
    k_final<<<(unsigned)B, 128, (size_t)((n + 3) & ~3) + 8 * (size_t)W, st>>>(fa);
    CK(cudaGetLastError());
    ctx->launches++;
    if (counters) {
        unsigned long long h[4];
        CK(cudaMemcpyAsync(h, w.counters, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 4; i++) counters[i] = (int64_t)h[i];
    }
    return 0;
}

// ------------------------------------------------------------------ measurement helpers -
static int time_probe(fbgnn_ctx *ctx, void (*kernel)(float *, int), int iters, double per_thread_ops, double *rate) {
    float *d = nullptr;
    CK(cudaMalloc(&d, 4));
    const int blocks = ctx->num_sms * 8, threads = 256;
    kernel<<<blocks, threads, 0, ctx->stream>>>(d, 16);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        best = std::min(best, ms);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaFree(d));
    *rate = per_thread_ops * (double)blocks * threads / (best * 1e-3);
    return 0;
}

extern "C" int fbgnn_sfu_peak(fbgnn_ctx *ctx, double *evals_per_s) {
    REQUIRE(ctx && evals_per_s, "NULL argument");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const int iters = 20000;
    return time_probe(ctx, k_sfu_peak, iters, 8.0 * iters, evals_per_s);
}

extern "C" int fbgnn_fma_peak(fbgnn_ctx *ctx, double *instr_per_s) {
    REQUIRE(ctx && instr_per_s, "NULL argument");
    if (set_device(ctx)) return FBGNN_E_CUDA;
    const int iters = 100000;
    return time_probe(ctx, k_fma_peak, iters, 8.0 * iters, instr_per_s);
}

extern "C" int fbgnn_math_probe(fbgnn_ctx *ctx, const char *fn, const float *x, float *y, int64_t n) {
    REQUIRE(ctx && fn && x && y && n >= 0, "bad argument");
    static const char *names[] = {"exp", "log", "log1p", "softplus", "phi4", "phi2", "tanh", "atanh"};
    int id = -1;
    for (int i = 0; i < 8; i++) if (!std::strcmp(fn, names[i])) id = i;
    REQUIRE(id >= 0, "unknown probe function '%s'", fn);
    if (set_device(ctx)) return FBGNN_E_CUDA;
    if (n == 0) return 0;
    k_math_probe<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(id, x, y, n);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}
