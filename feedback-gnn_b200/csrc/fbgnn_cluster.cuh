// fbgnn_cluster.cuh -- quaternary BP for codes whose per-frame state exceeds the shared memory of one SM
// (north_star: "a thread-block cluster for the larger codes").
//
// A thread-block cluster of C CTAs (2, 4 or 8 SMs) decodes ONE frame.  The variable nodes are cut into C contiguous
// ranges; CTA r keeps the messages of the edges of its variables (both sides, VN order, so a range of variables is a
// range of edges), their priors and their decisions in ITS shared memory.  The variable-node phase is local.  In the
// check-node phase a check gathers its incoming messages from whichever CTAs own them through distributed shared
// memory (cluster.map_shared_rank), updates them with the very code of the one-SM kernel, and scatters them back; the
// two phases are separated by cluster-wide barriers.  The epilogue's soft syndromes read the per-variable terms of
// remote CTAs the same way.  Same arithmetic, same orders as k_bp4: bit-exact with the oracle.
//
// Compared with the HBM-state fallback (k_bp4<..., GSTATE>): the 12 message accesses of a check go to a peer SM's
// shared memory (~215 cycles, no L2 / HBM traffic) instead of the L2; profiles/r02_cluster_vs_gstate.txt has the timing.
#pragma once
#include <cooperative_groups.h>

#include "fbgnn_kernels.cuh"

namespace fbgnn {
namespace cg = cooperative_groups;

constexpr int CL_MAX = 8;

struct ClusterPart {
    int C;                              // CTAs per cluster
    int v0[CL_MAX + 1];                 // variable ranges
    int ex0[CL_MAX + 1], ez0[CL_MAX + 1];   // first edge (VN order) of each range, per side
    int c0[CL_MAX + 1];                 // check ranges over the combined index space [0, m_x + m_z)
    int nv_max, ex_max, ez_max;         // largest range sizes: every CTA lays its shared memory out for these
};

__device__ __forceinline__ int cl_owner(const int *start, int C, int idx) {
    int r = 0;
#pragma unroll
    for (int i = 1; i < CL_MAX; i++) r += (i < C && idx >= start[i]) ? 1 : 0;
    return r;
}

__device__ const idx_t kIota64[64] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                      26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48,
                                      49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63};

// Dynamic shared memory of every CTA: float mx[ex_max], mz[ez_max], pri[(CONST_PRIOR ? 2 : 3) * nv_max];
//                                     u8 dec[nv_max]; int flag[CL_MAX]
//   DCMAX: bound on the check degree -- 8 keeps a check's gathered messages and their addresses in registers (the regular
//          product codes), 64 is the general case (local memory).
template <bool CONST_PRIOR, typename MATH, int DCMAX>
static __global__ void __launch_bounds__(512) k_bp4_cluster(const Bp4Args a, const ClusterPart P) {
    extern __shared__ float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const SideDev &X = a.X, &Z = a.Z;
    const int n = X.n, T = blockDim.x, tid = threadIdx.x;
    const int C = P.C, rank = (int)cluster.block_rank();
    const int64_t fi = blockIdx.x / C;
    const int64_t b = a.frame_list ? a.frame_list[fi] : fi;
    float *mx = smem, *mz = mx + P.ex_max, *pri = mz + P.ez_max;
    const int NV = P.nv_max;
    uint8_t *dec = (uint8_t *)(pri + (CONST_PRIOR ? 2 : 3) * NV);
    int *flag = (int *)(dec + ((NV + 3) & ~3));
    const int v0 = P.v0[rank], v1 = P.v0[rank + 1], ex0 = P.ex0[rank], ez0 = P.ez0[rank];
    const int c0 = P.c0[rank], c1 = P.c0[rank + 1];

    for (int e = tid; e < P.ex_max; e += T) mx[e] = 0.0f;
    for (int e = tid; e < P.ez_max; e += T) mz[e] = 0.0f;
    if (!CONST_PRIOR)
        for (int i = tid; i < 3 * (v1 - v0); i += T) {
            const int k = i / (v1 - v0), lv = i - k * (v1 - v0);
            pri[k * NV + lv] = a.llr(b, k, v0 + lv);
        }
    cluster.sync();

    // remote-capable pointer to the message of edge e (VN-order position) of one side
    auto msg_ptr = [&](bool isx, int e) -> float * {
        const int owner = cl_owner(isx ? P.ex0 : P.ez0, C, e);
        float *base = cluster.map_shared_rank(isx ? mx : mz, owner);
        return base + (e - (isx ? P.ex0[owner] : P.ez0[owner]));
    };

    for (int it = 0; it < a.num_iter; it++) {
        // variable nodes of this CTA's range (decoding_q.py:227-275): all local
        for (int v = v0 + tid; v < v1; v += T) {
            const int lv = v - v0;
            const float px = CONST_PRIOR ? a.prior : pri[lv];
            const float py = CONST_PRIOR ? a.prior : pri[NV + lv];
            const float pz = CONST_PRIOR ? a.prior : pri[2 * NV + lv];
            const int x0 = X.vn_ptr[v] - ex0, x1 = X.vn_ptr[v + 1] - ex0, z0 = Z.vn_ptr[v] - ez0, z1 = Z.vn_ptr[v + 1] - ez0;
            float Sx = 0.0f, Sz = 0.0f;
            for (int e = x0; e < x1; e++) Sx = FB_ADD(Sx, mx[e]);
            for (int e = z0; e < z1; e++) Sz = FB_ADD(Sz, mz[e]);
            const float ly = FB_ADD(FB_ADD(Sz, Sx), py);
            const float lx = FB_ADD(Sz, px);
            const float lz = FB_ADD(Sx, pz);
            const float num_hx = MATH::softplus(-lx), num_hz = MATH::softplus(-lz);
            if (MATH::kSharedVnCorr) {
                float ux, cx, uz, cz;
                vn_corr_pair<MATH, false>(lx, ly, lz, ux, cx, uz, cz);
                for (int e = x0; e < x1; e++) mx[e] = vn_msg(num_hx, mx[e], ux, cx);
                for (int e = z0; e < z1; e++) mz[e] = vn_msg(num_hz, mz[e], uz, cz);
                continue;
            }
            for (int e = x0; e < x1; e++) {
                const float m = mx[e];
                mx[e] = FB_SUB(num_hx, logaddexp_sat<MATH>(-FB_SUB(lz, m), -FB_SUB(ly, m)));
            }
            for (int e = z0; e < z1; e++) {
                const float m = mz[e];
                mz[e] = FB_SUB(num_hz, logaddexp_sat<MATH>(-FB_SUB(lx, m), -FB_SUB(ly, m)));
            }
        }
        cluster.sync();
        // check nodes of this CTA's share: gather over distributed shared memory, update, scatter
        for (int c = c0 + tid; c < c1; c += T) {
            const bool isx = c < X.m;
            const SideDev &S = isx ? X : Z;
            const int cc = isx ? c : c - X.m;
            const int k0 = S.cn_ptr[cc], deg = S.cn_ptr[cc + 1] - k0;
            float *ptr[DCMAX];
            float m[DCMAX];
#pragma unroll
            for (int k = 0; k < DCMAX; k++)
                if (k < deg) {
                    ptr[k] = msg_ptr(isx, S.cn_edge[k0 + k]);
                    m[k] = *ptr[k];
                }
            const int sb = isx ? a.sx(cc, b) : a.sz(cc, b);
            cn_update_one<true, MATH>(kIota64, 0, deg, m, sb, a.cn_type, a.factor);
#pragma unroll
            for (int k = 0; k < DCMAX; k++)
                if (k < deg) *ptr[k] = m[k];
        }
        cluster.sync();
    }

    // final messages (teacher-forced tests): every CTA writes its own edges
    if (a.msg_x.ptr) for (int e = tid; e < P.ex0[rank + 1] - ex0; e += T) a.msg_x(b, ex0 + e) = mx[e];
    if (a.msg_z.ptr) for (int e = tid; e < P.ez0[rank + 1] - ez0; e += T) a.msg_z(b, ez0 + e) = mz[e];

    // marginals, decision, per-variable terms of the soft syndromes (decoding_q.py:771-790, 455-464): local
    const bool want_logits = a.xl.ptr != nullptr || a.zl.ptr != nullptr;
    for (int v = v0 + tid; v < v1; v += T) {
        const int lv = v - v0;
        float Sx = 0.0f, Sz = 0.0f;
        for (int e = X.vn_ptr[v] - ex0; e < X.vn_ptr[v + 1] - ex0; e++) Sx = FB_ADD(Sx, mx[e]);
        for (int e = Z.vn_ptr[v] - ez0; e < Z.vn_ptr[v + 1] - ez0; e++) Sz = FB_ADD(Sz, mz[e]);
        const float px = CONST_PRIOR ? a.prior : pri[lv];
        const float py = CONST_PRIOR ? a.prior : pri[NV + lv];
        const float pz = CONST_PRIOR ? a.prior : pri[2 * NV + lv];
        const float ly = FB_ADD(FB_ADD(Sz, Sx), py);
        const float lx = FB_ADD(Sz, px);
        const float lz = FB_ADD(Sx, pz);
        if (a.Lx.ptr) a.Lx(b, v) = lx;
        if (a.Ly.ptr) a.Ly(b, v) = ly;
        if (a.Lz.ptr) a.Lz(b, v) = lz;
        int d = 0;
        float best = 0.0f;
        if (lx < best) { best = lx; d = 1; }
        if (lz < best) { best = lz; d = 2; }
        if (ly < best) { best = ly; d = 3; }
        if (want_logits) {
            const float llr_zp = FB_SUB(MATH::softplus(-lx), MATH::logaddexp(-lz, -ly));
            const float llr_xp = FB_SUB(MATH::softplus(-lz), MATH::logaddexp(-lx, -ly));
            d |= (llr_xp < 0.0f) << 2;
            d |= (llr_zp < 0.0f) << 3;
            pri[lv] = MATH::phi4(fabsf(llr_xp));
            pri[NV + lv] = MATH::phi4(fabsf(llr_zp));
        }
        dec[lv] = (uint8_t)d;
        if (a.xh.ptr) a.xh(b, v) = d & 1;
        if (a.zh.ptr) a.zh(b, v) = (d >> 1) & 1;
    }
    cluster.sync();

    // soft syndromes per check row and the syndrome match of the decision: per-variable terms over DSMEM
    int mismatch = 0;
    for (int c = c0 + tid; c < c1; c += T) {
        const bool isx = c < X.m;
        const SideDev &S = isx ? X : Z;
        const int cc = isx ? c : c - X.m;
        const int sbit = isx ? 3 : 2, dbit = isx ? 1 : 0;
        int par = 0, dpar = 0;
        float Tsum = 0.0f;
        for (int k = S.cn_ptr[cc]; k < S.cn_ptr[cc + 1]; k++) {
            const int v = S.cn_vn[k];
            const int owner = cl_owner(P.v0, C, v), lv = v - P.v0[owner];
            const int dv = cluster.map_shared_rank(dec, owner)[lv];
            par ^= (dv >> sbit) & 1;
            dpar ^= (dv >> dbit) & 1;
            if (want_logits) Tsum = FB_ADD(Tsum, cluster.map_shared_rank(pri, owner)[(isx ? NV : 0) + lv]);
        }
        mismatch |= dpar ^ (isx ? a.sx(cc, b) : a.sz(cc, b));
        if (want_logits) {
            float val = MATH::phi4(Tsum);
            val = par ? -val : val;
            if (isx) { if (a.zl.ptr) a.zl(cc, b) = val; }
            else     { if (a.xl.ptr) a.xl(cc, b) = val; }
        }
    }
    if (a.vbits) {
        mismatch = __syncthreads_or(mismatch);
        if (tid == 0) cluster.map_shared_rank(flag, 0)[rank] = mismatch;       // every rank reports to rank 0
        const int act_in = a.active_in ? a.active_in[b] : 1;
        if (act_in) {
            uint8_t *vb = a.vbits + b * n;
            for (int v = v0 + tid; v < v1; v += T) vb[v] = (vb[v] & 3) | ((dec[v - v0] & 3) << 2);
        }
        cluster.sync();
        if (rank == 0 && tid == 0) {
            int any = 0;
            for (int r = 0; r < C; r++) any |= flag[r];
            const int act_out = act_in && any;
            a.active_out[b] = (uint8_t)act_out;
            if (a.rounds && act_out) a.rounds[b] += 1;
            if (a.next_list && act_out) a.next_list[atomicAdd(a.next_count, 1)] = (int)b;
        }
        if (a.stats && rank == 0 && tid == 0) { atomicAdd(a.stats, 1ull); atomicAdd(a.stats + 1, (unsigned long long)a.num_iter); }
    } else {
        cluster.sync();                 // no CTA may leave while a peer still reads its shared memory
        if (a.stats && rank == 0 && tid == 0) { atomicAdd(a.stats, 1ull); atomicAdd(a.stats + 1, (unsigned long long)a.num_iter); }
    }
}

}  // namespace fbgnn
