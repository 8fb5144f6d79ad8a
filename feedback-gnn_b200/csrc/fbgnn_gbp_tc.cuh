// fbgnn_gbp_tc.cuh -- tensor-core (tcgen05 / TMEM) path of the GNN_BP4 node updates, sm_100a.
//
// After the factoring of fbgnn_kernels.cuh (k_gbp_*_f) every matrix product of GNN_BP4 is a per-NODE
// product [rows = frames x nodes] x [K <= 40] x [N <= 80] with the weights as the stationary operand --
// the one place on this path where "frames x nodes x hidden" is a real dense GEMM.
// This file runs those products on the 5th-generation tensor cores:
//
//   * one group of 128 threads owns a tile of 128 rows; thread r <-> row r <-> TMEM lane r, so the
//     per-row work between the products (edge loop with tanh, bias, mean) stays thread-local;
//   * the A operand (the row's activations) goes registers -> TMEM with tcgen05.st, never through shared
//     memory; the B operands (weights) are staged once per CTA in shared memory in the canonical K-major
//     no-swizzle UMMA layout; D accumulates in TMEM and comes back with tcgen05.ld;
//   * float32 accuracy from TF32 hardware by the three-product split  A B ~= Ah Bh + Al Bh + Ah Bl
//     (Ah, Bh = operands rounded to TF32, Al, Bl = remainders): measured max error 1e-6 relative to the
//     largest output (tools/micro/umma_test.cu), i.e. float32 rounding level;
//   * a CTA holds two independent groups (TMEM columns [0,128) and [128,256)) that share the staged
//     weights and synchronise separately (named barriers + one mbarrier each); two CTAs per SM use all 512
//     TMEM columns.
//
// The results agree with the FMA path / the oracle to float32 re-association accuracy, not bit for bit
// (tests/test_gnn_bp4.py states the tolerance); the FMA path remains the default.
#ifndef FBGNN_GBP_TC_CUH
#define FBGNN_GBP_TC_CUH

#include "fbgnn_kernels.cuh"

namespace fbgnn {
namespace tc {

// ------------------------------------------------------------------ tcgen05 primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 8 consecutive 32-bit columns of the calling thread's TMEM lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t v[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __uint_as_float(r[q]);
}

// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, M = 128; issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

// Canonical K-major no-swizzle tile of an [NPAD x KPAD] operand of 32-bit elements: 8 x 16 B core matrices,
// adjacent along K (leading byte offset 128), 8-row groups KPAD / 4 * 128 bytes apart (stride byte offset).
__host__ __device__ inline int b_tile_offset(int n, int k, int kpad) {
    return (n >> 3) * (kpad / 4) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr, int kpad) {
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((kpad / 4) * 128) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);     // version 1, no swizzle
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = npad
__host__ __device__ inline uint32_t idesc_tf32(int npad) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((128u >> 4) << 24);
}
// x = hi + lo with hi representable in TF32 (10 explicit mantissa bits, round to nearest)
__host__ __device__ inline float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
    uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r;
#endif
}

// ------------------------------------------------------------------ packed weights
// Each operand: [hi tile][lo tile], KPAD x NPAD floats each.  Offsets in floats.
constexpr int t(int k, int n) { return 2 * k * n; }
struct VnW {       // update_h_vn (+ sender halves of update_h_cn's message MLPs)
    static constexpr int B1X = 0, B1Z = B1X + t(24, 48);               // receiver halves, K = D (24), N = H (48)
    static constexpr int W2X = B1Z + t(24, 48), W2Z = W2X + t(40, 32); // message output layers, K = H, N = M (32)
    static constexpr int W3AB = W2Z + t(40, 32);                       // embed MLP layer 1, rows of [m_x, m_z]: K = 2M, N = H (48)
    static constexpr int W3C = W3AB + t(40, 48);                       //                    rows of h_vn:       K = D (24)
    static constexpr int W4 = W3C + t(24, 48);                         // embed MLP layer 2: K = H, N = D (32)
    static constexpr int W5 = W4 + t(40, 32);                          // sender halves for the CN update: K = D (24), N = 2H (80)
    static constexpr int BIAS = W5 + t(24, 80);                        // b1x[40] b1z[40] b2x[20] b2z[20] b3[40] b4[20]
    static constexpr int total = BIAS + 180;
};
struct CnW {       // update_h_cn of ONE side (+ sender half of update_h_vn's message MLP of that side)
    static constexpr int B1 = 0, W2 = B1 + t(24, 48);
    static constexpr int W3A = W2 + t(40, 32);                         // rows [m (20), logit (1), 0, 0, 0]: K = 24
    static constexpr int W3B = W3A + t(24, 48);                        // rows of h_cn: K = D (24)
    static constexpr int W4 = W3B + t(24, 48);
    static constexpr int W5 = W4 + t(40, 32);                          // K = D (24), N = H (48)
    static constexpr int BIAS = W5 + t(24, 48);                        // b1[40] b2[20] b3[40] b4[20]
    static constexpr int total = BIAS + 120;
};

// ------------------------------------------------------------------ group helpers
struct Grp {
    uint32_t tmem;        // TMEM address of the group's column 0, lane 0
    uint32_t lane_base;   // + this warp's lane quarter
    uint32_t bar;         // shared-memory address of the group's mbarrier
    uint32_t phase;
    int gid;
    bool leader;
};

__device__ __forceinline__ void grp_sync(int gid) { asm volatile("bar.sync %0, 128;" :: "r"(gid + 1) : "memory"); }

// registers -> TMEM as the A operand: hi parts at columns [col_hi, col_hi + KPAD), remainders at col_lo
template <int KPAD>
__device__ __forceinline__ void store_a(const Grp &g, int col_hi, int col_lo, const float *a) {
    __syncwarp();
#pragma unroll
    for (int k0 = 0; k0 < KPAD; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float h = tf32_hi(a[k0 + q]);
            hi[q] = __float_as_uint(h);
            lo[q] = __float_as_uint(a[k0 + q] - h);
        }
        tmem_st8(g.lane_base + col_hi + k0, hi);
        tmem_st8(g.lane_base + col_lo + k0, lo);
    }
}

// D[128 x npad] (at col_d) = or += A[128 x kpad] * B, all rows of the group; returns when D is readable
__device__ __forceinline__ void gemm(Grp &g, int col_ahi, int col_alo, int kpad, int col_d, uint32_t b_saddr, int npad,
                                     bool accumulate) {
    __syncwarp();
    tmem_wait_st();
    tmem_wait_ld();
    tc_fence_before();
    grp_sync(g.gid);
    if (g.leader) {
        tc_fence_after();
        const uint32_t idesc = idesc_tf32(npad);
        const uint32_t bl = b_saddr + (uint32_t)(npad * kpad * 4);
        for (int s = 0; s < kpad / 8; s++) {
            const uint64_t dh = b_desc(b_saddr + s * 256, kpad), dl = b_desc(bl + s * 256, kpad);
            umma_tf32_ts(g.tmem + col_d, g.tmem + col_ahi + s * 8, dh, idesc, (accumulate || s > 0) ? 1u : 0u);
            umma_tf32_ts(g.tmem + col_d, g.tmem + col_alo + s * 8, dh, idesc, 1u);
            umma_tf32_ts(g.tmem + col_d, g.tmem + col_ahi + s * 8, dl, idesc, 1u);
        }
        umma_commit(g.bar);
    }
    mbar_wait(g.bar, g.phase);
    g.phase ^= 1u;
    tc_fence_after();
    __syncwarp();
}

// Edge phase of one receiver (see gbp_recv_factored): the receiver halves come from TMEM columns
// [col_base, col_base + 40), the signed sums of hidden activations go to TMEM as the next A operand
// (hi at [0, 40), lo at [40, 80)).  Returns the sum of the message signs.
template <typename MATH, typename SENDER>
__device__ __forceinline__ int recv_tc(const Grp &g, int col_base, const float *__restrict__ b1, int e0, int e1,
                                       const float *__restrict__ rows, int row_stride, uint32_t *codes, int cstride,
                                       SENDER sender) {
    constexpr int JB = 8, H = 40;
    gbp_stage_codes(codes, cstride, e0, e1, sender);
    auto code_at = [&](int e) -> uint32_t { return (e - e0 < GBP_CODE_CAP) ? codes[(e - e0) * cstride] : sender(e); };
    int ssum = 0;
    uint32_t c0 = 0;
    float4 nx0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), nx1 = nx0;
    if (e0 < e1) {
        c0 = code_at(e0);
        const float *pf = rows + (int64_t)(c0 >> 1) * row_stride;
        nx0 = *reinterpret_cast<const float4 *>(pf); nx1 = *reinterpret_cast<const float4 *>(pf + 4);
    }
#pragma unroll 1
    for (int j = 0; j < H; j += JB) {
        float bs[JB], hs[JB];
        __syncwarp();
        tmem_ld8(g.lane_base + col_base + j, bs);
        const float4 ba = *reinterpret_cast<const float4 *>(b1 + j), bb = *reinterpret_cast<const float4 *>(b1 + j + 4);
        bs[0] += ba.x; bs[1] += ba.y; bs[2] += ba.z; bs[3] += ba.w; bs[4] += bb.x; bs[5] += bb.y; bs[6] += bb.z; bs[7] += bb.w;
#pragma unroll
        for (int q = 0; q < JB; q++) hs[q] = 0.0f;
        uint32_t cc = c0;
#pragma unroll 1
        for (int e = e0; e < e1; e++) {
            const float pq[JB] = {nx0.x, nx0.y, nx0.z, nx0.w, nx1.x, nx1.y, nx1.z, nx1.w};
            const bool neg = (cc & 1u) != 0;
            if (e + 1 < e1) {
                cc = code_at(e + 1);
                const float *pf = rows + (int64_t)(cc >> 1) * row_stride + j;
                nx0 = *reinterpret_cast<const float4 *>(pf); nx1 = *reinterpret_cast<const float4 *>(pf + 4);
            }
            if (j == 0) ssum += neg ? -1 : 1;
#pragma unroll
            for (int q = 0; q < JB; q++) {
                float t = MATH::tanh(pq[q] + bs[q]);
                hs[q] += neg ? -t : t;
            }
        }
        if (j + JB < H && e0 < e1) {
            const float *pf = rows + (int64_t)(c0 >> 1) * row_stride + j + JB;
            nx0 = *reinterpret_cast<const float4 *>(pf); nx1 = *reinterpret_cast<const float4 *>(pf + 4);
        }
        uint32_t hi[JB], lo[JB];
#pragma unroll
        for (int q = 0; q < JB; q++) {
            const float h = tf32_hi(hs[q]);
            hi[q] = __float_as_uint(h);
            lo[q] = __float_as_uint(hs[q] - h);
        }
        __syncwarp();
        tmem_st8(g.lane_base + j, hi);
        tmem_st8(g.lane_base + H + j, lo);
    }
    return ssum;
}

// The packed operand tiles (57 - 90 KB) come into shared memory with bulk asynchronous copies (TMA, 1-D form):
// one thread posts the transfers against an mbarrier, nobody spends load / store instructions on them, and the
// tensor core later reads them through the same (async) proxy that wrote them.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ Grp cta_setup(float *sm, const float *__restrict__ wsrc, int nfloats, uint64_t *bars, uint32_t *tslot) {
    __shared__ uint64_t load_bar;
    if (threadIdx.x < 32) tmem_alloc(tslot, 256);
    if (threadIdx.x == 32) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&load_bar, 1);
        fence_async_smem();
        const uint32_t bytes = (uint32_t)nfloats * 4u, bar = smem_u32(&load_bar), dst = smem_u32(sm);
        mbar_expect_tx(bar, bytes);
        constexpr uint32_t CHUNK = 32768;                        // multiples of 16 bytes
        for (uint32_t off = 0; off < bytes; off += CHUNK)
            bulk_g2s(dst + off, reinterpret_cast<const char *>(wsrc) + off, min(CHUNK, bytes - off), bar);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(smem_u32(&load_bar), 0);
    Grp g;
    g.gid = threadIdx.x >> 7;
    const int t = threadIdx.x & 127;
    g.tmem = *tslot + (uint32_t)(g.gid * 128);
    g.lane_base = g.tmem + ((uint32_t)((t >> 5) * 32) << 16);
    g.bar = smem_u32(&bars[g.gid]);
    g.phase = 0;
    g.leader = t == 0;
    return g;
}

__device__ __forceinline__ void cta_teardown(const uint32_t *tslot) {
    tmem_wait_ld();
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(*tslot, 256);
}

// ------------------------------------------------------------------ UpdateVNEmbeddings, tensor-core form
// TMEM columns of a group:  A operands at [0, 80) (hi, lo), D at [80, 128); the last product swaps the roles.
template <typename MATH>
__global__ void __launch_bounds__(256, 2) k_gbp_vn_tc(const GbpArgs a, const float *__restrict__ wtc) {
    constexpr int D = 20, H = 40, M = 20;
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    Grp g = cta_setup(sm, wtc, VnW::total, bars, &tslot);
    const uint32_t sbase = smem_u32(sm);
    const float *bias = sm + VnW::BIAS;
    uint32_t *codes = reinterpret_cast<uint32_t *>(sm + VnW::total) + threadIdx.x;      // [GBP_CODE_CAP][256]
    const int n = a.X.n, mt = a.X.m + a.Z.m, t = threadIdx.x & 127;
    const int64_t total = a.B * n, ntiles = (total + 127) / 128;
    for (int64_t tile = (int64_t)blockIdx.x * 2 + g.gid; tile < ntiles; tile += (int64_t)gridDim.x * 2) {
        const int64_t it = tile * 128 + t;
        const bool valid = it < total;
        const int64_t b = valid ? it / n : 0;
        const int v = valid ? (int)(it - b * n) : 0;
        float own[24], mm[2 * M];
#pragma unroll
        for (int k = 0; k < 24; k++) own[k] = 0.0f;
        if (valid) {
#pragma unroll
            for (int k = 0; k < D; k += 4) {
                const float4 v4 = *reinterpret_cast<const float4 *>(a.h_vn + it * D + k);
                own[k] = v4.x; own[k + 1] = v4.y; own[k + 2] = v4.z; own[k + 3] = v4.w;
            }
        }
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
            const SideDev &S = side ? a.Z : a.X;
            const float *pf_b = a.pfc + (b * mt + (side ? a.X.m : 0)) * H;
            const uint8_t *sy = side ? a.sz + b * a.Z.m : a.sx + b * a.X.m;
            const idx_t *vn_cn = S.vn_cn;
            const int e0 = valid ? S.vn_ptr[v] : 0, e1 = valid ? S.vn_ptr[v + 1] : 0;
            store_a<24>(g, 0, 24, own);
            gemm(g, 0, 24, 24, 80, sbase + 4 * (side ? VnW::B1Z : VnW::B1X), 48, false);
            const int ssum = recv_tc<MATH>(g, 80, bias + side * H, e0, e1, pf_b, H, codes, 256,
                                           [&](int e) { const uint32_t c = vn_cn[e]; return (c << 1) | (sy[c] != 0 ? 1u : 0u); });
            gemm(g, 0, 40, 40, 80, sbase + 4 * (side ? VnW::W2Z : VnW::W2X), 32, false);
            const float fs = (float)ssum, dg = (float)(e1 - e0);
            const float *b2 = bias + 2 * H + side * M;
#pragma unroll
            for (int c0 = 0; c0 < 24; c0 += 8) {
                float d[8];
                tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (c0 + q < M) {
                        float r = d[q] + fs * b2[c0 + q];
                        if (a.reduce == 0 && e1 > e0) r = r / dg;
                        if (side == 0) mm[c0 + q] = r;
                        else mm[M + c0 + q] = r;
                    }
                }
            }
        }
        store_a<40>(g, 0, 40, mm);
        gemm(g, 0, 40, 40, 80, sbase + 4 * VnW::W3AB, 48, false);
        store_a<24>(g, 0, 24, own);
        gemm(g, 0, 24, 24, 80, sbase + 4 * VnW::W3C, 48, true);
        float hid[H];
        const float *b3 = bias + 2 * H + 2 * M;
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 8) {
            float d[8];
            tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
            for (int q = 0; q < 8; q++) hid[c0 + q] = MATH::tanh(d[q] + b3[c0 + q]);
        }
        store_a<40>(g, 0, 40, hid);
        gemm(g, 0, 40, 40, 80, sbase + 4 * VnW::W4, 32, false);
        float out[24];
        const float *b4 = b3 + H;
#pragma unroll
        for (int c0 = 0; c0 < 24; c0 += 8) {
            float d[8];
            tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
            for (int q = 0; q < 8; q++) out[c0 + q] = (c0 + q < D) ? d[q] + b4[c0 + q] : 0.0f;
        }
        if (valid) {
#pragma unroll
            for (int k = 0; k < D; k += 4)
                *reinterpret_cast<float4 *>(a.h_vn + it * D + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
        }
        store_a<24>(g, 80, 104, out);
        gemm(g, 80, 104, 24, 0, sbase + 4 * VnW::W5, 80, false);
#pragma unroll 1
        for (int c0 = 0; c0 < 2 * H; c0 += 8) {
            float d[8];
            __syncwarp();
            tmem_ld8(g.lane_base + c0, d);
            if (valid) {
                float *dst = a.pfv + it * 2 * H + c0;
                *reinterpret_cast<float4 *>(dst) = make_float4(d[0], d[1], d[2], d[3]);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(d[4], d[5], d[6], d[7]);
            }
        }
    }
    cta_teardown(&tslot);
}

// ------------------------------------------------------------------ UpdateCNEmbeddings (one side), tensor-core form
template <typename MATH>
__global__ void __launch_bounds__(256, 2) k_gbp_cn_tc(const GbpArgs a, const float *__restrict__ wtc, int side) {
    constexpr int D = 20, H = 40, M = 20;
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    Grp g = cta_setup(sm, wtc, CnW::total, bars, &tslot);
    const uint32_t sbase = smem_u32(sm);
    const float *bias = sm + CnW::BIAS;
    uint32_t *codes = reinterpret_cast<uint32_t *>(sm + CnW::total) + threadIdx.x;      // [GBP_CODE_CAP][256]
    const SideDev &S = side ? a.Z : a.X;
    const int n = a.X.n, mt = a.X.m + a.Z.m, ms = S.m, coff = side ? a.X.m : 0, t = threadIdx.x & 127;
    const idx_t *cn_vn = S.cn_vn;
    float *hcs = side ? a.hcz : a.hcx;
    const uint8_t *sys = side ? a.sz : a.sx;
    const int64_t total = a.B * ms, ntiles = (total + 127) / 128;
    for (int64_t tile = (int64_t)blockIdx.x * 2 + g.gid; tile < ntiles; tile += (int64_t)gridDim.x * 2) {
        const int64_t it = tile * 128 + t;
        const bool valid = it < total;
        const int64_t b = valid ? it / ms : 0;
        const int c = valid ? (int)(it - b * ms) : 0;
        float own[24], in[24];
#pragma unroll
        for (int k = 0; k < 24; k++) { own[k] = 0.0f; in[k] = 0.0f; }
        int e0 = 0, e1 = 0;
        if (valid) {
#pragma unroll
            for (int k = 0; k < D; k += 4) {
                const float4 v4 = *reinterpret_cast<const float4 *>(hcs + it * D + k);
                own[k] = v4.x; own[k + 1] = v4.y; own[k + 2] = v4.z; own[k + 3] = v4.w;
            }
            e0 = S.cn_ptr[c]; e1 = S.cn_ptr[c + 1];
            if (!a.zero_logits) {
                const float lgv = a.lg[b * mt + coff + c];
                in[M] = sys[it] ? -lgv : lgv;                                            // logit * (1 - 2 s)
            }
        }
        const float *pf_b = a.pfv + b * n * 2 * H + (side ? H : 0);
        store_a<24>(g, 0, 24, own);
        gemm(g, 0, 24, 24, 80, sbase + 4 * CnW::B1, 48, false);
        recv_tc<MATH>(g, 80, bias, e0, e1, pf_b, 2 * H, codes, 256, [&](int e) { return (uint32_t)cn_vn[e] << 1; });
        gemm(g, 0, 40, 40, 80, sbase + 4 * CnW::W2, 32, false);
        const float dg = (float)(e1 - e0);
        const float *b2 = bias + H;
#pragma unroll
        for (int c0 = 0; c0 < 24; c0 += 8) {
            float d[8];
            tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (c0 + q < M) {
                    float r = d[q] + dg * b2[c0 + q];
                    if (a.reduce == 0 && e1 > e0) r = r / dg;
                    in[c0 + q] = r;
                }
            }
        }
        store_a<24>(g, 0, 24, in);
        gemm(g, 0, 24, 24, 80, sbase + 4 * CnW::W3A, 48, false);
        store_a<24>(g, 0, 24, own);
        gemm(g, 0, 24, 24, 80, sbase + 4 * CnW::W3B, 48, true);
        float hid[H];
        const float *b3 = bias + H + M;
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 8) {
            float d[8];
            tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
            for (int q = 0; q < 8; q++) hid[c0 + q] = MATH::tanh(d[q] + b3[c0 + q]);
        }
        store_a<40>(g, 0, 40, hid);
        gemm(g, 0, 40, 40, 80, sbase + 4 * CnW::W4, 32, false);
        float out[24];
        const float *b4 = b3 + H;
#pragma unroll
        for (int c0 = 0; c0 < 24; c0 += 8) {
            float d[8];
            tmem_ld8(g.lane_base + 80 + c0, d);
#pragma unroll
            for (int q = 0; q < 8; q++) out[c0 + q] = (c0 + q < D) ? d[q] + b4[c0 + q] : 0.0f;
        }
        if (valid) {
#pragma unroll
            for (int k = 0; k < D; k += 4)
                *reinterpret_cast<float4 *>(hcs + it * D + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
        }
        store_a<24>(g, 80, 104, out);
        gemm(g, 80, 104, 24, 0, sbase + 4 * CnW::W5, 48, false);
#pragma unroll 1
        for (int c0 = 0; c0 < H; c0 += 8) {
            float d[8];
            __syncwarp();
            tmem_ld8(g.lane_base + c0, d);
            if (valid) {
                float *dst = a.pfc + (b * mt + coff + c) * H + c0;
                *reinterpret_cast<float4 *>(dst) = make_float4(d[0], d[1], d[2], d[3]);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(d[4], d[5], d[6], d[7]);
            }
        }
    }
    cta_teardown(&tslot);
}

}  // namespace tc
}  // namespace fbgnn

#endif  // FBGNN_GBP_TC_CUH
