/* fbgnn_oracle.c -- CPU ORACLE for the BP -> feedback-GNN -> BP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (feedback-gnn_b200/) calls, links or
 * imports this file; it is used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / --impl reference legs as the CHECKER and the CPU baseline.
 *
 * It is a plain-C restatement of the reference's algorithm, function by function:
 *   orc_bp4      QLDPCBPDecoder.call            sionna/fec/ldpc/decoding_q.py:661-797
 *                  _vn_update                   decoding_q.py:227-275
 *                  _cn_update_phi / _phi        decoding_q.py:365-431
 *                  _cn_update_tanh              decoding_q.py:313-363
 *                  _cn_update_minsum            decoding_q.py:539-644
 *                  cal_logit/_cn_update_phi_loss decoding_q.py:433-471
 *   orc_bp2      LDPCBPDecoder.call (syndrome)  sionna/fec/ldpc/decoding.py:875-1048
 *                  _vn_update 511-535, _cn_update_phi 635-690, _phi 625-633
 *   orc_gnn      Feedback_GNN.call              sionna/fec/ldpc/feedback_gnn.py:161-188, gnn.py:31-69
 *   orc_pauli    Pauli.call (non-wt)            sionna/channel/pauli.py:98-108
 *   orc_pipeline Sandwich_BP_GNN_Evaluation_Model.call   feedback_gnn.py:293-361
 *   orc_bsc_pipeline BP_BSC_Model.call          feedback_gnn.py:207-229
 *
 * PARITY STATUS: "parity unpinned" by a live run of the reference -- TensorFlow is not
 * installable in this image, so the reference cannot be executed.  The
 * oracle is pinned instead by the reference's stored notebook outputs: the deterministic
 * known answers (phi values, first-stage marginal extrema of examples/n1270.ipynb cell 12)
 * and the logical-error-rate tables (binomial confidence intervals); see tests/.
 *
 * Arithmetic: float32 throughout, every elementary function taken from
 * feedback-gnn_b200/csrc/fb_math.h (the arithmetic specification shared with the CUDA
 * kernels, which makes GPU-vs-oracle comparisons bit-exact).  An independent numpy
 * restatement (oracle/np_oracle.py, numpy's own exp/log/tanh) cross-checks that header.
 *
 * Edge ordering (the reference leaves it to an unstable argsort, decoding_q.py:63-85):
 * VN order = edges sorted by (vn, cn); CN order = edges sorted by (cn, vn); all sums and
 * products run sequentially in that order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fb_math.h"
#include "fb_umma.h"

#define ORC_CN_PHI    0
#define ORC_CN_TANH   1
#define ORC_CN_MINSUM 2

typedef struct {
    int32_t n, m, E;
    const int32_t *vn_ptr;   /* [n+1] edge ranges per VN, edges sorted by (vn, cn)          */
    const int32_t *vn_cn;    /* [E]   check index of each edge (VN order)                   */
    const int32_t *cn_ptr;   /* [m+1] edge ranges per CN, edges sorted by (cn, vn)          */
    const int32_t *cn_edge;  /* [E]   position in VN order of each edge listed in CN order  */
    const int32_t *cn_vn;    /* [E]   variable index of each edge listed in CN order        */
} orc_side_t;

typedef struct {             /* sparse rows of a binary matrix (perp matrices, logicals)    */
    int32_t m;
    const int32_t *ptr;      /* [m+1] */
    const int32_t *col;      /* [nnz] increasing within a row */
} orc_rows_t;

typedef struct {             /* Feedback_GNN weights, Keras get_weights() order             */
    int32_t H, M;            /* hidden units, message dims                                  */
    int32_t act;             /* 0 tanh, 1 relu, 2 linear                                    */
    int32_t reduce;          /* 0 mean, 1 sum, 2 max, 3 min                                 */
    const float *W0, *b0;    /* [H,3], [3]     _llr_inv_embed                               */
    const float *W1x, *b1x;  /* [4,H], [H]     vn_msg_mlp_x layer 0                         */
    const float *W2x, *b2x;  /* [H,M], [M]     vn_msg_mlp_x layer 1                         */
    const float *W1z, *b1z, *W2z, *b2z;
    const float *W3, *b3;    /* [2M+3,H], [H]  vn_embed_mlp                                 */
    int32_t gemm;            /* 0: FP32 FMAs; 1: the tensor-core form (fbgnn_gnn_tc.cuh), see gnn_frame_tc */
} orc_gnn_t;

/* ---------------------------------------------------------------- Philox4x32-10 ---- */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11). */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    philox4x32_10(ctr, key, out);
}

/* uniform in [0,1) of qubit q of global frame f: word (q & 3) of Philox(ctr = {f_lo, f_hi,
 * q >> 2, stream}, key = {seed_lo, seed_hi}), top 24 bits scaled by 2^-24. */
static float frame_uniform(uint64_t seed, uint64_t frame, uint32_t q, uint32_t stream) {
    uint32_t ctr[4] = { (uint32_t)frame, (uint32_t)(frame >> 32), q >> 2, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t r[4];
    philox4x32_10(ctr, key, r);
    return (float)(r[q & 3] >> 8) * 5.9604644775390625e-08f;
}

/* Pauli.call, non-wt branch (pauli.py:98-108).  thr = {px, px - py, (px + pz) - py} in f32. */
static void pauli_frame(uint64_t seed, uint64_t frame, int n, const float thr[3],
                        uint8_t *nx, uint8_t *nz) {
    for (int q = 0; q < n; q++) {
        float u = frame_uniform(seed, frame, (uint32_t)q, 0u);
        nx[q] = (uint8_t)(u < thr[0]);
        nz[q] = (uint8_t)((u >= thr[1]) && (u < thr[2]));
    }
}

/* Pauli.call, wt branch (pauli.py:80-96): `wt` distinct positions per frame (the reference takes
 * the first wt entries of a shuffle; here a partial Fisher-Yates driven by Philox stream 2), then one
 * uniform per position (stream 3): u < 2/3 sets the X bit, u > 1/3 the Z bit. */
static void pauli_wt_frame(uint64_t seed, uint64_t frame, int n, int wt, uint8_t *nx, uint8_t *nz,
                           uint16_t *idx) {
    for (int q = 0; q < n; q++) { idx[q] = (uint16_t)q; nx[q] = 0; nz[q] = 0; }
    if (wt > n) wt = n;
    for (int i = 0; i < wt; i++) {
        float u = frame_uniform(seed, frame, (uint32_t)i, 2u);
        int j = i + (int)FB_MUL(u, (float)(n - i));
        uint16_t t = idx[i]; idx[i] = idx[j]; idx[j] = t;
        float w = frame_uniform(seed, frame, (uint32_t)i, 3u);
        nx[idx[i]] = (uint8_t)(w < 0.6666667f);
        nz[idx[i]] = (uint8_t)(w > 0.33333334f);
    }
}

void orc_pauli_wt(uint64_t seed, uint64_t first_frame, int64_t B, int n, int wt, uint8_t *noise_x,
                  uint8_t *noise_z) {
#pragma omp parallel
    {
        uint16_t *idx = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; b++)
            pauli_wt_frame(seed, first_frame + (uint64_t)b, n, wt, noise_x + b * n, noise_z + b * n, idx);
        free(idx);
    }
}

void orc_pauli(uint64_t seed, uint64_t first_frame, int64_t B, int n, const float *thr,
               uint8_t *noise_x, uint8_t *noise_z) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        pauli_frame(seed, first_frame + (uint64_t)b, n, thr, noise_x + b * n, noise_z + b * n);
}

/* BinarySymmetricChannel (discrete_channel.py:385-396): a Bernoulli(p) flip per bit.  The
 * reference draws it with a Gumbel-softmax trick; statistically it is u < p. */
static void bsc_frame(uint64_t seed, uint64_t frame, int n, float p, uint8_t *noise) {
    for (int q = 0; q < n; q++)
        noise[q] = (uint8_t)(frame_uniform(seed, frame, (uint32_t)q, 1u) < p);
}

void orc_bsc(uint64_t seed, uint64_t first_frame, int64_t B, int n, float p, uint8_t *noise) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        bsc_frame(seed, first_frame + (uint64_t)b, n, p, noise + b * n);
}

/* syndrome of one frame: s[c] = XOR over the check's variables (int_mod_2(H @ e)) */
static void syndrome_frame(const orc_side_t *S, const uint8_t *e, uint8_t *s) {
    for (int c = 0; c < S->m; c++) {
        uint8_t a = 0;
        for (int k = S->cn_ptr[c]; k < S->cn_ptr[c + 1]; k++) a ^= e[S->cn_vn[k]];
        s[c] = a;
    }
}

static int rows_any_parity(const orc_rows_t *R, const uint8_t *e) {
    for (int r = 0; r < R->m; r++) {
        uint8_t a = 0;
        for (int k = R->ptr[r]; k < R->ptr[r + 1]; k++) a ^= e[R->col[k]];
        if (a) return 1;
    }
    return 0;
}

/* ---------------------------------------------------------------- check nodes ------ */
/* One side, all checks.  msg holds v2c on entry and c2v on exit (VN order, in place).
 * phi4 != 0 selects the quaternary decoder's phi, else the binary decoder's. */
static void cn_update(const orc_side_t *S, int cn_type, int phi4, float factor,
                      const uint8_t *synd, float *msg, float *work) {
    for (int c = 0; c < S->m; c++) {
        const int k0 = S->cn_ptr[c], k1 = S->cn_ptr[c + 1];
        const float ssign = (synd && synd[c]) ? -1.0f : 1.0f;
        if (cn_type == ORC_CN_PHI) {
            /* decoding_q.py:392-429 */
            float sgn = 1.0f, T = 0.0f;
            for (int k = k0; k < k1; k++) {
                float m = msg[S->cn_edge[k]];
                if (m < 0.0f) sgn = -sgn;
                float a = phi4 ? fb_m_phi4f(fabsf(m)) : fb_m_phi2f(fabsf(m));
                work[k - k0] = a;
                T = FB_ADD(T, a);
            }
            sgn = sgn * ssign;
            for (int k = k0; k < k1; k++) {
                int e = S->cn_edge[k];
                float s = (msg[e] < 0.0f) ? -sgn : sgn;
                float x = FB_SUB(T, work[k - k0]);
                float v = phi4 ? fb_m_phi4f(x) : fb_m_phi2f(x);
                msg[e] = FB_MUL(FB_MUL(s, v), factor);
            }
        } else if (cn_type == ORC_CN_TANH) {
            /* decoding_q.py:327-361 */
            float P = 1.0f;
            for (int k = k0; k < k1; k++) {
                float t = fb_m_tanhf(FB_MUL(msg[S->cn_edge[k]], 0.5f));
                if (t == 0.0f) t = 1e-12f;
                work[k - k0] = t;
                P = FB_MUL(P, t);
            }
            P = FB_MUL(P, ssign);
            for (int k = k0; k < k1; k++) {
                float v = FB_MUL(FB_DIV(1.0f, work[k - k0]), P);
                if (fabsf(v) < 1e-7f) v = 0.0f;
                v = FB_FMIN(FB_FMAX(v, -FB_ATANH_CLIP), FB_ATANH_CLIP);
                v = FB_MUL(2.0f, fb_m_atanhf(v));
                msg[S->cn_edge[k]] = FB_MUL(v, factor);
            }
        } else {
            /* min-sum, decoding_q.py:551-642 */
            const float LARGE = 10000.0f;
            float sgn = 1.0f, mn = INFINITY;
            for (int k = k0; k < k1; k++) {
                float m = msg[S->cn_edge[k]];
                m = FB_FMIN(FB_FMAX(m, -FB_LLR_MAX), FB_LLR_MAX);
                if (m < 0.0f) sgn = -sgn;
                float a = fabsf(m);
                work[k - k0] = a;
                if (a < mn) mn = a;
            }
            sgn = sgn * ssign;
            float mn2 = INFINITY, sum = 0.0f;
            for (int k = k0; k < k1; k++) {
                float d = FB_SUB(work[k - k0], mn);
                if (d == 0.0f) d = LARGE;
                work[k - k0] = d;
                if (d < mn2) mn2 = d;
                sum = FB_ADD(sum, d);
            }
            mn2 = FB_ADD(mn2, mn);
            float node_sum = FB_SUB(sum, 2.0f * LARGE - 1.0f);
            float sg = (node_sum > 0.0f) ? 1.0f : ((node_sum < 0.0f) ? -1.0f : 0.0f);
            float dm = FB_MUL(0.5f, FB_SUB(1.0f, sg));
            float mne = FB_ADD(FB_MUL(FB_SUB(1.0f, dm), mn), FB_MUL(dm, mn2));
            for (int k = k0; k < k1; k++) {
                int e = S->cn_edge[k];
                float m = msg[e];
                float s = (m < 0.0f) ? -sgn : sgn;
                float v = (work[k - k0] == LARGE) ? mne : mn;
                msg[e] = FB_MUL(FB_MUL(s, v), factor);
            }
        }
    }
}

static int max_cn_degree(const orc_side_t *S) {
    int d = 1;
    for (int c = 0; c < S->m; c++) {
        int k = S->cn_ptr[c + 1] - S->cn_ptr[c];
        if (k > d) d = k;
    }
    return d;
}

/* ---------------------------------------------------------------- quaternary BP ---- */
/* marginals of one frame (decoding_q.py:244-251) */
static void bp4_marginals(const orc_side_t *X, const orc_side_t *Z, const float *mx,
                          const float *mz, const float *llrx, const float *llry,
                          const float *llrz, float *Lx, float *Ly, float *Lz) {
    for (int v = 0; v < X->n; v++) {
        float Sx = 0.0f, Sz = 0.0f;
        for (int e = X->vn_ptr[v]; e < X->vn_ptr[v + 1]; e++) Sx = FB_ADD(Sx, mx[e]);
        for (int e = Z->vn_ptr[v]; e < Z->vn_ptr[v + 1]; e++) Sz = FB_ADD(Sz, mz[e]);
        Ly[v] = FB_ADD(FB_ADD(Sz, Sx), llry[v]);
        Lx[v] = FB_ADD(Sz, llrx[v]);
        Lz[v] = FB_ADD(Sx, llrz[v]);
    }
}

/* soft syndromes of one frame (cal_logit, decoding_q.py:455-471).  rows_x: rows of
 * _pcm_x_perp evaluated on llr_x'; rows_z: rows of _pcm_z_perp evaluated on llr_z'. */
static void bp4_logits(int n, const orc_rows_t *rows_x, const orc_rows_t *rows_z,
                       const float *Lx, const float *Ly, const float *Lz,
                       float *x_logit, float *z_logit, float *wa, float *wb) {
    /* wa = llr_x', wb = llr_z' */
    for (int v = 0; v < n; v++) {
        wb[v] = FB_SUB(fb_m_softplusf(-Lx[v]), fb_m_logaddexpf(-Lz[v], -Ly[v]));
        wa[v] = FB_SUB(fb_m_softplusf(-Lz[v]), fb_m_logaddexpf(-Lx[v], -Ly[v]));
    }
    for (int side = 0; side < 2; side++) {
        const orc_rows_t *R = side ? rows_z : rows_x;
        const float *l = side ? wb : wa;
        float *out = side ? z_logit : x_logit;
        if (!R || !out) continue;
        for (int r = 0; r < R->m; r++) {
            float sgn = 1.0f, T = 0.0f;
            for (int k = R->ptr[r]; k < R->ptr[r + 1]; k++) {
                float m = l[R->col[k]];
                if (m < 0.0f) sgn = -sgn;
                T = FB_ADD(T, fb_m_phi4f(fabsf(m)));
            }
            out[r] = FB_MUL(sgn, fb_m_phi4f(T));
        }
    }
}

/* One frame of QLDPCBPDecoder.call.  mx/mz: work [E_x]/[E_z]; on exit they hold the c2v
 * messages of the last iteration (VN order). */
typedef struct {               /* optional per-iteration soft syndromes (stage_two / trainable,      */
    const orc_rows_t *rows_x;  /* decoding_q.py:743-746, 779-781): slot 2*it = x_logit, 2*it+1 = z_logit */
    const orc_rows_t *rows_z;
    float *out;                /* [2*num_iter+2][m][B] */
    int64_t B, b;
    int m;
    float *wa;                 /* work [2n] */
} orc_iterlog_t;

static void iterlog_write(const orc_iterlog_t *il, int n, int slot, const float *Lx, const float *Ly,
                          const float *Lz, float *xl, float *zl) {
    bp4_logits(n, il->rows_x, il->rows_z, Lx, Ly, Lz, xl, zl, il->wa, il->wa + n);
    for (int r = 0; r < il->rows_x->m; r++) il->out[((int64_t)(2 * slot) * il->m + r) * il->B + il->b] = xl[r];
    for (int r = 0; r < il->rows_z->m; r++) il->out[((int64_t)(2 * slot + 1) * il->m + r) * il->B + il->b] = zl[r];
}

static void bp4_frame_il(const orc_side_t *X, const orc_side_t *Z, int cn_type, int num_iter,
                         float factor, const float *llrx, const float *llry, const float *llrz,
                         const uint8_t *sx, const uint8_t *sz, float *mx, float *mz,
                         float *Lx, float *Ly, float *Lz, uint8_t *xh, uint8_t *zh, float *work,
                         const orc_iterlog_t *il, float *xl, float *zl, int *iters_used) {
    const int n = X->n;
    memset(mx, 0, sizeof(float) * (size_t)X->E);
    memset(mz, 0, sizeof(float) * (size_t)Z->E);
    if (iters_used) *iters_used = num_iter;
    for (int it = 0; it < num_iter; it++) {
        if (il) {              /* marginals the VN update of this iteration is about to compute */
            bp4_marginals(X, Z, mx, mz, llrx, llry, llrz, Lx, Ly, Lz);
            iterlog_write(il, n, it, Lx, Ly, Lz, xl, zl);
        }
        /* VN update, decoding_q.py:227-275 */
        for (int v = 0; v < n; v++) {
            float Sx = 0.0f, Sz = 0.0f;
            for (int e = X->vn_ptr[v]; e < X->vn_ptr[v + 1]; e++) Sx = FB_ADD(Sx, mx[e]);
            for (int e = Z->vn_ptr[v]; e < Z->vn_ptr[v + 1]; e++) Sz = FB_ADD(Sz, mz[e]);
            float ly = FB_ADD(FB_ADD(Sz, Sx), llry[v]);
            float lx = FB_ADD(Sz, llrx[v]);
            float lz = FB_ADD(Sx, llrz[v]);
            float num_hx = fb_m_softplusf(-lx);
            float num_hz = fb_m_softplusf(-lz);
            if (fb_math_mode) {         /* SFU arithmetic: one correction term per side (fb_math.h fb_sfu_vn_corr) */
                const float ux = FB_FMIN(lz, ly), cx = fb_sfu_vn_corr(lz, ly), uz = FB_FMIN(lx, ly), cz = fb_sfu_vn_corr(lx, ly);
                for (int e = X->vn_ptr[v]; e < X->vn_ptr[v + 1]; e++) mx[e] = fb_sfu_vn_msg(num_hx, mx[e], ux, cx);
                for (int e = Z->vn_ptr[v]; e < Z->vn_ptr[v + 1]; e++) mz[e] = fb_sfu_vn_msg(num_hz, mz[e], uz, cz);
                continue;
            }
            for (int e = X->vn_ptr[v]; e < X->vn_ptr[v + 1]; e++) {
                float a = FB_SUB(lz, mx[e]), b = FB_SUB(ly, mx[e]);
                mx[e] = FB_SUB(num_hx, fb_m_logaddexpf(-a, -b));
            }
            for (int e = Z->vn_ptr[v]; e < Z->vn_ptr[v + 1]; e++) {
                float a = FB_SUB(lx, mz[e]), b = FB_SUB(ly, mz[e]);
                mz[e] = FB_SUB(num_hz, fb_m_logaddexpf(-a, -b));
            }
        }
        cn_update(X, cn_type, 1, factor, sx, mx, work);
        cn_update(Z, cn_type, 1, factor, sz, mz, work);
        if (iters_used) {
            /* opt-in early stop (NOT what the reference does, which always runs num_iter iterations):
             * leave the loop as soon as the hard decision of the current messages reproduces the syndrome.  The frame's
             * outputs are then exactly those of a decoder configured with num_iter = *iters_used. */
            bp4_marginals(X, Z, mx, mz, llrx, llry, llrz, Lx, Ly, Lz);
            for (int v = 0; v < n; v++) {
                int d = 0;
                float best = 0.0f;
                if (Lx[v] < best) { best = Lx[v]; d = 1; }
                if (Lz[v] < best) { best = Lz[v]; d = 2; }
                if (Ly[v] < best) { best = Ly[v]; d = 3; }
                xh[v] = (uint8_t)(d & 1);
                zh[v] = (uint8_t)(d >> 1);
            }
            int bad = 0;
            for (int c = 0; c < X->m && !bad; c++) {
                int par = sx[c];
                for (int k = X->cn_ptr[c]; k < X->cn_ptr[c + 1]; k++) par ^= zh[X->cn_vn[k]];
                bad |= par;
            }
            for (int c = 0; c < Z->m && !bad; c++) {
                int par = sz[c];
                for (int k = Z->cn_ptr[c]; k < Z->cn_ptr[c + 1]; k++) par ^= xh[Z->cn_vn[k]];
                bad |= par;
            }
            if (!bad) { *iters_used = it + 1; num_iter = it + 1; break; }
        }
    }
    bp4_marginals(X, Z, mx, mz, llrx, llry, llrz, Lx, Ly, Lz);
    if (il) iterlog_write(il, n, num_iter, Lx, Ly, Lz, xl, zl);
    /* argmin over [0, Lx, Lz, Ly], first minimum wins (decoding_q.py:786-790) */
    for (int v = 0; v < n; v++) {
        int d = 0;
        float best = 0.0f;
        if (Lx[v] < best) { best = Lx[v]; d = 1; }
        if (Lz[v] < best) { best = Lz[v]; d = 2; }
        if (Ly[v] < best) { best = Ly[v]; d = 3; }
        xh[v] = (uint8_t)(d & 1);
        zh[v] = (uint8_t)(d >> 1);
    }
}

static int g_pipeline_early_stop = 0;      /* set per orc_pipeline call (read-only inside the parallel region) */

static void bp4_frame(const orc_side_t *X, const orc_side_t *Z, int cn_type, int num_iter,
                      float factor, const float *llrx, const float *llry, const float *llrz,
                      const uint8_t *sx, const uint8_t *sz, float *mx, float *mz,
                      float *Lx, float *Ly, float *Lz, uint8_t *xh, uint8_t *zh, float *work) {
    int used;
    bp4_frame_il(X, Z, cn_type, num_iter, factor, llrx, llry, llrz, sx, sz, mx, mz, Lx, Ly, Lz, xh, zh,
                 work, NULL, NULL, NULL, g_pipeline_early_stop ? &used : NULL);
}

/* Batched QLDPCBPDecoder.call with the reference's tensor layouts:
 *   llr [B,3,n] (or NULL: the scalar `prior` is used for all three), synd_x [m_x,B],
 *   synd_z [m_z,B] (uint8 0/1), outputs Lx,Ly,Lz [B,n] f32, x_hat,z_hat [B,n] u8,
 *   x_logit [rows_x.m, B], z_logit [rows_z.m, B] (may be NULL), msg_x [B,E_x], msg_z
 *   [B,E_z] final c2v messages (may be NULL). */
void orc_bp4(const orc_side_t *X, const orc_side_t *Z, const orc_rows_t *rows_x,
             const orc_rows_t *rows_z, int cn_type, int num_iter, float factor, int64_t B,
             const float *llr, float prior, const uint8_t *synd_x, const uint8_t *synd_z,
             float *Lx, float *Ly, float *Lz, uint8_t *x_hat, uint8_t *z_hat,
             float *x_logit, float *z_logit, float *msg_x, float *msg_z, float *llr_hat, uint8_t *iters_out) {
    /* iters_out != NULL: opt-in early stop, iters_out[b] = iterations executed for frame b */
    const int n = X->n, mxn = X->m, mzn = Z->m;
    int wd = max_cn_degree(X), wz = max_cn_degree(Z);
    if (wz > wd) wd = wz;
#pragma omp parallel
    {
        float *mx = (float *)malloc(sizeof(float) * (size_t)(X->E + 1));
        float *mz = (float *)malloc(sizeof(float) * (size_t)(Z->E + 1));
        float *pri = (float *)malloc(sizeof(float) * 3 * (size_t)n);
        float *work = (float *)malloc(sizeof(float) * (size_t)wd);
        float *wa = (float *)malloc(sizeof(float) * 2 * (size_t)n);
        uint8_t *sx = (uint8_t *)malloc((size_t)mxn + 1), *sz = (uint8_t *)malloc((size_t)mzn + 1);
        float *xl = (float *)malloc(sizeof(float) * (size_t)((rows_x ? rows_x->m : 0) + 1));
        float *zl = (float *)malloc(sizeof(float) * (size_t)((rows_z ? rows_z->m : 0) + 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; b++) {
            for (int v = 0; v < 3 * n; v++) pri[v] = llr ? llr[b * 3 * n + v] : prior;
            for (int c = 0; c < mxn; c++) sx[c] = synd_x[(int64_t)c * B + b];
            for (int c = 0; c < mzn; c++) sz[c] = synd_z[(int64_t)c * B + b];
            orc_iterlog_t il = { rows_x, rows_z, llr_hat, B, b, rows_x ? rows_x->m : 0, wa };
            int used = num_iter;
            bp4_frame_il(X, Z, cn_type, num_iter, factor, pri, pri + n, pri + 2 * n, sx, sz, mx, mz,
                         Lx + b * n, Ly + b * n, Lz + b * n, x_hat + b * n, z_hat + b * n, work,
                         llr_hat ? &il : NULL, xl, zl, iters_out ? &used : NULL);
            if (iters_out) iters_out[b] = (uint8_t)used;
            if (x_logit || z_logit) {
                bp4_logits(n, x_logit ? rows_x : NULL, z_logit ? rows_z : NULL, Lx + b * n,
                           Ly + b * n, Lz + b * n, xl, zl, wa, wa + n);
                if (x_logit) for (int r = 0; r < rows_x->m; r++) x_logit[(int64_t)r * B + b] = xl[r];
                if (z_logit) for (int r = 0; r < rows_z->m; r++) z_logit[(int64_t)r * B + b] = zl[r];
            }
            if (msg_x) memcpy(msg_x + b * X->E, mx, sizeof(float) * (size_t)X->E);
            if (msg_z) memcpy(msg_z + b * Z->E, mz, sizeof(float) * (size_t)Z->E);
        }
        free(mx); free(mz); free(pri); free(work); free(wa); free(sx); free(sz); free(xl); free(zl);
    }
}

/* ---------------------------------------------------------------- binary BP -------- */
/* One frame of LDPCBPDecoder.call with is_syndrome (decoding.py:905-1034).  `logit` is the
 * layer input (log p1/p0); returns the soft output (logits) in `soft` and the hard decision
 * (1 if logit > 0) in `hard`. */
/* edge_w (optional): the trainable decoder's per-edge weights on the v2c messages, VN order (decoding.py:981-983);
 * msg_in / msg_out (optional): the stateful decoder's c2v messages before / after the call (decoding.py:947-953). */
static void bp2_frame_ex(const orc_side_t *S, int cn_type, int num_iter, float factor,
                         const float *logit, const uint8_t *synd, float *msg, float *soft,
                         uint8_t *hard, float *work, float *llr, const float *edge_w, const float *msg_in,
                         float *msg_out) {
    const int n = S->n;
    for (int v = 0; v < n; v++) {
        float l = FB_FMIN(FB_FMAX(logit[v], -FB_LLR_MAX), FB_LLR_MAX);   /* decoding.py:918-920 */
        llr[v] = -l;                                                   /* decoding.py:940 */
    }
    if (msg_in) memcpy(msg, msg_in, sizeof(float) * (size_t)S->E);
    else memset(msg, 0, sizeof(float) * (size_t)S->E);
    for (int it = 0; it < num_iter; it++) {
        for (int v = 0; v < n; v++) {                                  /* decoding.py:511-535 */
            float s = 0.0f;
            for (int e = S->vn_ptr[v]; e < S->vn_ptr[v + 1]; e++) s = FB_ADD(s, msg[e]);
            s = FB_ADD(s, llr[v]);
            for (int e = S->vn_ptr[v]; e < S->vn_ptr[v + 1]; e++) {
                msg[e] = FB_SUB(s, msg[e]);
                if (edge_w) msg[e] = FB_MUL(msg[e], edge_w[e]);
            }
        }
        cn_update(S, cn_type, 0, factor, synd, msg, work);
    }
    if (msg_out) memcpy(msg_out, msg, sizeof(float) * (size_t)S->E);
    for (int v = 0; v < n; v++) {                                      /* decoding.py:1026-1034 */
        float s = 0.0f;
        for (int e = S->vn_ptr[v]; e < S->vn_ptr[v + 1]; e++) s = FB_ADD(s, msg[e]);
        float x = -FB_ADD(llr[v], s);
        soft[v] = x;
        hard[v] = (uint8_t)(0.0f < x);
    }
}

/* llr [B,n] logits, synd [m,B] or NULL, soft [B,n], hard [B,n] */
static void bp2_frame(const orc_side_t *S, int cn_type, int num_iter, float factor,
                      const float *logit, const uint8_t *synd, float *msg, float *soft,
                      uint8_t *hard, float *work, float *llr) {
    bp2_frame_ex(S, cn_type, num_iter, factor, logit, synd, msg, soft, hard, work, llr, NULL, NULL, NULL);
}

void orc_bp2(const orc_side_t *S, int cn_type, int num_iter, float factor, int64_t B,
             const float *llr, const uint8_t *synd, float *soft, uint8_t *hard, const float *edge_w,
             const float *msg_in, float *msg_out) {
    const int n = S->n, m = S->m;
    int wd = max_cn_degree(S);
#pragma omp parallel
    {
        float *msg = (float *)malloc(sizeof(float) * (size_t)(S->E + 1));
        float *work = (float *)malloc(sizeof(float) * (size_t)wd);
        float *l = (float *)malloc(sizeof(float) * (size_t)n);
        uint8_t *s = (uint8_t *)malloc((size_t)m + 1);
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; b++) {
            if (synd) for (int c = 0; c < m; c++) s[c] = synd[(int64_t)c * B + b];
            bp2_frame_ex(S, cn_type, num_iter, factor, llr + b * n, synd ? s : NULL, msg,
                         soft + b * n, hard + b * n, work, l, edge_w, msg_in ? msg_in + b * S->E : NULL,
                         msg_out ? msg_out + b * S->E : NULL);
        }
        free(msg); free(work); free(l); free(s);
    }
}

/* ---------------------------------------------------------------- feedback GNN ----- */
static float gnn_act(int act, float x) {
    if (act == 0) return fb_m_tanhf(x);
    if (act == 1) return x > 0.0f ? x : 0.0f;
    return x;
}

/* hidden layer of the edge MLP for edge e of variable v: features [h_cn[c], Lx, Ly, Lz]
 * (feedback_gnn.py:175-178); the three per-variable terms are accumulated first, the check-node
 * term last */
static float gnn_hidden(const orc_gnn_t *G, const float *W1, const float *b1, float hc, const float f3[3], int j) {
    const int H = G->H;
    float a = 0.0f;
    for (int k = 0; k < 3; k++) a = FB_FMA(f3[k], W1[(k + 1) * H + j], a);
    a = FB_FMA(hc, W1[j], a);
    if (b1) a = FB_ADD(a, b1[j]);
    return gnn_act(G->act, a);
}

/* messages of one side into one variable node, reduced (feedback_gnn.py:175-184).
 * reduce_op "mean" / "sum": the output layer of the edge MLP is linear, so it is applied ONCE to the
 * sum of the hidden activations over the node's edges --
 *     sum_e (W2^T t_e + b2) = W2^T (sum_e t_e) + deg * b2        (mean: divided by deg) --
 * which is the reference's value up to float32 re-association (np_oracle.py keeps the literal
 * per-edge form; tests/test_oracle.py bounds the difference).  "max" / "min" keep the per-edge form. */
static void gnn_side(const orc_side_t *S, const orc_gnn_t *G, const float *W1, const float *b1,
                     const float *W2, const float *b2, const float *h_cn, int v,
                     const float f3[3], float *red, float *hid, float *msg) {
    const int H = G->H, M = G->M;
    const int e0 = S->vn_ptr[v], e1 = S->vn_ptr[v + 1];
    if (e1 == e0) { for (int i = 0; i < M; i++) red[i] = 0.0f; return; }
    if (G->reduce <= 1) {
        const float deg = (float)(e1 - e0);
        for (int j = 0; j < H; j++) {
            float hs = gnn_hidden(G, W1, b1, h_cn[S->vn_cn[e0]], f3, j);
            for (int e = e0 + 1; e < e1; e++) hs = FB_ADD(hs, gnn_hidden(G, W1, b1, h_cn[S->vn_cn[e]], f3, j));
            hid[j] = hs;
        }
        for (int i = 0; i < M; i++) {
            float a = 0.0f;
            for (int j = 0; j < H; j++) a = FB_FMA(hid[j], W2[j * M + i], a);
            if (G->reduce == 0) {
                a = FB_DIV(a, deg);
                if (b2) a = FB_ADD(a, b2[i]);
            } else if (b2) {
                a = FB_FMA(deg, b2[i], a);
            }
            red[i] = a;
        }
        return;
    }
    for (int e = e0; e < e1; e++) {
        const float hc = h_cn[S->vn_cn[e]];
        for (int j = 0; j < H; j++) hid[j] = gnn_hidden(G, W1, b1, hc, f3, j);
        for (int i = 0; i < M; i++) {
            float a = 0.0f;
            for (int j = 0; j < H; j++) a = FB_FMA(hid[j], W2[j * M + i], a);
            if (b2) a = FB_ADD(a, b2[i]);
            msg[i] = a;
        }
        for (int i = 0; i < M; i++) {
            if (e == e0) red[i] = msg[i];
            else if (G->reduce == 2) red[i] = (msg[i] > red[i]) ? msg[i] : red[i];
            else red[i] = (msg[i] < red[i]) ? msg[i] : red[i];
        }
    }
}

static void gnn_frame_tc(const orc_side_t *X, const orc_side_t *Z, const orc_gnn_t *G,
                         const float *Lx, const float *Ly, const float *Lz, const float *logit_hx,
                         const float *logit_hz, const uint8_t *sx, const uint8_t *sz,
                         float *ox, float *oy, float *oz);

/* One frame of Feedback_GNN.call.  logit_hx [m_x] pairs with hx rows, logit_hz [m_z] with hz
 * rows (the caller passes the decoder's z_logit and x_logit, feedback_gnn.py:335). */
static void gnn_frame(const orc_side_t *X, const orc_side_t *Z, const orc_gnn_t *G,
                      const float *Lx, const float *Ly, const float *Lz, const float *logit_hx,
                      const float *logit_hz, const uint8_t *sx, const uint8_t *sz,
                      float *ox, float *oy, float *oz, float *work) {
    const int n = X->n, H = G->H, M = G->M;
    if (G->gemm == 1) { gnn_frame_tc(X, Z, G, Lx, Ly, Lz, logit_hx, logit_hz, sx, sz, ox, oy, oz); return; }
    float *hcx = work, *hcz = hcx + X->m, *in = hcz + Z->m, *hid = in + 2 * M + 3, *msg = hid + H;
    for (int c = 0; c < X->m; c++) hcx[c] = FB_MUL(logit_hx[c], sx[c] ? -1.0f : 1.0f);
    for (int c = 0; c < Z->m; c++) hcz[c] = FB_MUL(logit_hz[c], sz[c] ? -1.0f : 1.0f);
    for (int v = 0; v < n; v++) {
        const float f3[3] = { Lx[v], Ly[v], Lz[v] };
        gnn_side(X, G, G->W1x, G->b1x, G->W2x, G->b2x, hcx, v, f3, in, hid, msg);
        gnn_side(Z, G, G->W1z, G->b1z, G->W2z, G->b2z, hcz, v, f3, in + M, hid, msg);
        in[2 * M] = f3[0]; in[2 * M + 1] = f3[1]; in[2 * M + 2] = f3[2];
        for (int j = 0; j < H; j++) {
            float a = 0.0f;
            for (int k = 0; k < 2 * M + 3; k++) a = FB_FMA(in[k], G->W3[k * H + j], a);
            if (G->b3) a = FB_ADD(a, G->b3[j]);
            hid[j] = gnn_act(G->act, a);
        }
        float o[3];
        for (int i = 0; i < 3; i++) {
            float a = 0.0f;
            for (int j = 0; j < H; j++) a = FB_FMA(hid[j], G->W0[j * 3 + i], a);
            if (G->b0) a = FB_ADD(a, G->b0[i]);
            o[i] = a;
        }
        ox[v] = o[0]; oy[v] = o[1]; oz[v] = o[2];
    }
}

/* The tensor-core form of Feedback_GNN.call, operation by operation as csrc/fbgnn_gnn_tc.cuh evaluates it: the first
 * layers of the edge MLPs, tanh and the 3-wide output layer in float32 FMAs in the kernel's order; the three dense
 * products (hidden sums x W2x / W2z, messages x W3[0:2M]) as tcgen05.mma kind::tf32 steps on TF32 hi / lo splits,
 * emulated exactly by fb_umma.h.  tanh, reduce_op mean / sum, H and 2M multiples of 8. */
static void gnn_frame_tc(const orc_side_t *X, const orc_side_t *Z, const orc_gnn_t *G,
                         const float *Lx, const float *Ly, const float *Lz, const float *logit_hx,
                         const float *logit_hz, const uint8_t *sx, const uint8_t *sz,
                         float *ox, float *oy, float *oz) {
    const int n = X->n, H = G->H, M = G->M;
    float *w2h[2], *w2l[2];
    float *w3h = (float *)malloc(sizeof(float) * 2 * (size_t)(2 * M) * H), *w3l = w3h + (size_t)(2 * M) * H;
    for (int sd = 0; sd < 2; sd++) {
        const float *W2 = sd ? G->W2z : G->W2x;
        w2h[sd] = (float *)malloc(sizeof(float) * 2 * (size_t)H * M); w2l[sd] = w2h[sd] + (size_t)H * M;
        /* mean: the kernel folds 1 / degree into the W2 tiles as a float32 factor (regular degree, taken from variable 0) */
        const orc_side_t *S0 = sd ? Z : X;
        const float inv = (G->reduce == 0) ? FB_DIV(1.0f, (float)(S0->vn_ptr[1] - S0->vn_ptr[0])) : 1.0f;
        for (int i = 0; i < H * M; i++) {
            const float w = FB_MUL(W2[i], inv);
            w2h[sd][i] = fb_tf32_hi(w); w2l[sd][i] = FB_SUB(w, w2h[sd][i]);
        }
    }
    for (int i = 0; i < 2 * M * H; i++) { w3h[i] = fb_tf32_hi(G->W3[i]); w3l[i] = FB_SUB(G->W3[i], w3h[i]); }
    float *hi = (float *)malloc(sizeof(float) * 2 * (size_t)(H > 2 * M ? H : 2 * M)), *lo = hi + (H > 2 * M ? H : 2 * M);
    float *mm = (float *)malloc(sizeof(float) * 2 * (size_t)M);
    for (int v = 0; v < n; v++) {
        const float f1 = Lx[v], f2 = Ly[v], f3 = Lz[v];
        for (int sd = 0; sd < 2; sd++) {
            const orc_side_t *S = sd ? Z : X;
            const float *W1 = sd ? G->W1z : G->W1x, *b1 = sd ? G->b1z : G->b1x, *b2 = sd ? G->b2z : G->b2x;
            const float *logit = sd ? logit_hz : logit_hx;
            const uint8_t *sy = sd ? sz : sx;
            const int e0 = S->vn_ptr[v], e1 = S->vn_ptr[v + 1];
            const float dg = (float)(e1 - e0);
            for (int j = 0; j < H; j++) {
                const float base = FB_FMA(f3, W1[3 * H + j], FB_FMA(f2, W1[2 * H + j], FB_FMA(f1, W1[H + j], b1 ? b1[j] : 0.0f)));
                float hs = 0.0f;
                for (int e = e0; e < e1; e++) {
                    const int c = S->vn_cn[e];
                    const float hc = sy[c] ? -logit[c] : logit[c];
                    const float t = gnn_act(G->act, FB_FMA(hc, W1[j], base));
                    hs = (e == e0) ? t : FB_ADD(hs, t);
                }
                hi[j] = fb_tf32_hi(hs); lo[j] = FB_SUB(hs, hi[j]);
            }
            for (int i = 0; i < M; i++) {
                const float d = fb_umma_dot3(hi, lo, w2h[sd] + i, w2l[sd] + i, M, H);
                const float bb = b2 ? b2[i] : 0.0f;
                mm[sd * M + i] = (G->reduce == 0) ? FB_ADD(d, bb) : FB_FMA(dg, bb, d);
            }
        }
        for (int k = 0; k < 2 * M; k++) { hi[k] = fb_tf32_hi(mm[k]); lo[k] = FB_SUB(mm[k], hi[k]); }
        float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
        for (int j = 0; j < H; j++) {
            const float d = fb_umma_dot3(hi, lo, w3h + j, w3l + j, H, 2 * M);
            float h = FB_FMA(f1, G->W3[(2 * M) * H + j], d);
            h = FB_FMA(f2, G->W3[(2 * M + 1) * H + j], h);
            h = FB_FMA(f3, G->W3[(2 * M + 2) * H + j], h);
            h = gnn_act(G->act, FB_ADD(h, G->b3 ? G->b3[j] : 0.0f));
            o0 = FB_FMA(h, G->W0[j * 3 + 0], o0);
            o1 = FB_FMA(h, G->W0[j * 3 + 1], o1);
            o2 = FB_FMA(h, G->W0[j * 3 + 2], o2);
        }
        ox[v] = FB_ADD(o0, G->b0 ? G->b0[0] : 0.0f);
        oy[v] = FB_ADD(o1, G->b0 ? G->b0[1] : 0.0f);
        oz[v] = FB_ADD(o2, G->b0 ? G->b0[2] : 0.0f);
    }
    free(w2h[0]); free(w2h[1]); free(w3h); free(hi); free(mm);
}

static size_t gnn_work_floats(const orc_side_t *X, const orc_side_t *Z, const orc_gnn_t *G) {
    return (size_t)X->m + (size_t)Z->m + 2 * (size_t)G->M + 3 + (size_t)G->H + (size_t)G->M + 8;
}

/* h_vn [B,n,3], logit_hx [m_x,B], logit_hz [m_z,B], synd_* [m,B], out [B,n,3] */
void orc_gnn(const orc_side_t *X, const orc_side_t *Z, const orc_gnn_t *G, int64_t B,
             const float *h_vn, const float *logit_hx, const float *logit_hz,
             const uint8_t *synd_x, const uint8_t *synd_z, float *out) {
    const int n = X->n;
#pragma omp parallel
    {
        float *work = (float *)malloc(sizeof(float) * gnn_work_floats(X, Z, G));
        float *L = (float *)malloc(sizeof(float) * 6 * (size_t)n);
        float *lx = (float *)malloc(sizeof(float) * (size_t)(X->m + Z->m + 2));
        uint8_t *s = (uint8_t *)malloc((size_t)(X->m + Z->m + 2));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; b++) {
            for (int v = 0; v < n; v++)
                for (int k = 0; k < 3; k++) L[k * n + v] = h_vn[(b * n + v) * 3 + k];
            for (int c = 0; c < X->m; c++) { lx[c] = logit_hx[(int64_t)c * B + b]; s[c] = synd_x[(int64_t)c * B + b]; }
            for (int c = 0; c < Z->m; c++) { lx[X->m + c] = logit_hz[(int64_t)c * B + b]; s[X->m + c] = synd_z[(int64_t)c * B + b]; }
            gnn_frame(X, Z, G, L, L + n, L + 2 * n, lx, lx + X->m, s, s + X->m,
                      L + 3 * n, L + 4 * n, L + 5 * n, work);
            for (int v = 0; v < n; v++)
                for (int k = 0; k < 3; k++) out[(b * n + v) * 3 + k] = L[(3 + k) * n + v];
        }
        free(work); free(L); free(lx); free(s);
    }
}


/* ---------------------------------------------------------------- feedback GNN, any MLP depth ------- */
/* Feedback_GNN with num_mlp_layers = L != 2 (feedback_gnn.py:110-127: the edge MLPs have L-1 hidden layers of H units
 * and a linear output of M, the embedding MLP has L-1 hidden layers, _llr_inv_embed maps its output -- or, for L = 1,
 * the concatenated input itself -- to 3).  The two-layer case keeps its own (factored) arithmetic above; this general
 * form evaluates every edge's MLP in full and reduces the messages in edge order.
 * Packed weights: dense layers one after the other, each [K_in x K_out] row-major followed by K_out biases (zeros when
 * use_bias is off): _llr_inv_embed, vn_msg_mlp_x (L layers), vn_msg_mlp_z (L layers), vn_embed_mlp (L-1 layers). */
typedef struct {
    int32_t H, M, L, act, reduce, use_bias;
    const float *w;
} orc_gnnd_t;

#define GNND_MAX 128
static const float *gnnd_dense(const float *w, int kin, int kout, int use_bias, int act, const float *in, float *out) {
    for (int j = 0; j < kout; j++) {
        float a = 0.0f;
        for (int k = 0; k < kin; k++) a = FB_FMA(in[k], w[k * kout + j], a);
        if (use_bias) a = FB_ADD(a, w[kin * kout + j]);
        out[j] = gnn_act(act, a);
    }
    return w + kin * kout + kout;
}

static void gnnd_vn(const orc_side_t *X, const orc_side_t *Z, const orc_gnnd_t *G, const float *hcx, const float *hcz,
                    int v, const float f3[3], float out[3]) {
    const int H = G->H, M = G->M, L = G->L;
    float in[2 * GNND_MAX + 3], a[GNND_MAX], c[GNND_MAX];
    const float *w_inv = G->w;
    const float *w = w_inv + ((L == 1 ? 2 * M + 3 : H) * 3 + 3);
    for (int side = 0; side < 2; side++) {
        const orc_side_t *S = side ? Z : X;
        const float *hc = side ? hcz : hcx;
        float *red = in + side * M;
        const int e0 = S->vn_ptr[v], e1 = S->vn_ptr[v + 1];
        const float *wend = w;
        for (int i = 0; i < M; i++) red[i] = 0.0f;
        for (int e = e0; e < e1; e++) {
            a[0] = hc[S->vn_cn[e]]; a[1] = f3[0]; a[2] = f3[1]; a[3] = f3[2];
            const float *wl = w;
            int kin = 4;
            float *src = a, *dst = c;
            for (int l = 0; l < L; l++) {
                const int kout = (l == L - 1) ? M : H;
                wl = gnnd_dense(wl, kin, kout, G->use_bias, (l == L - 1) ? 2 : G->act, src, dst);
                float *t = src; src = dst; dst = t;
                kin = kout;
            }
            wend = wl;
            for (int i = 0; i < M; i++) {
                if (e == e0) red[i] = src[i];
                else if (G->reduce <= 1) red[i] = FB_ADD(red[i], src[i]);
                else if (G->reduce == 2) red[i] = (src[i] > red[i]) ? src[i] : red[i];
                else red[i] = (src[i] < red[i]) ? src[i] : red[i];
            }
        }
        if (e1 == e0) {                  /* skip over this side's layers */
            int kin = 4;
            for (int l = 0; l < L; l++) { const int kout = (l == L - 1) ? M : H; wend += kin * kout + kout; kin = kout; }
        } else if (G->reduce == 0) {
            const float deg = (float)(e1 - e0);
            for (int i = 0; i < M; i++) red[i] = FB_DIV(red[i], deg);
        }
        w = wend;
    }
    in[2 * M] = f3[0]; in[2 * M + 1] = f3[1]; in[2 * M + 2] = f3[2];
    float *src = in, *dst = a;
    int kin = 2 * M + 3;
    for (int l = 0; l < L - 1; l++) {
        w = gnnd_dense(w, kin, H, G->use_bias, G->act, src, dst);
        src = dst; dst = (dst == a) ? c : a;
        kin = H;
    }
    gnnd_dense(w_inv, kin, 3, G->use_bias, 2, src, out);
}

void orc_gnn_deep(const orc_side_t *X, const orc_side_t *Z, const orc_gnnd_t *G, int64_t B,
                  const float *h_vn, const float *logit_hx, const float *logit_hz,
                  const uint8_t *synd_x, const uint8_t *synd_z, float *out) {
    const int n = X->n;
#pragma omp parallel
    {
        float *hc = (float *)malloc(sizeof(float) * (size_t)(X->m + Z->m + 2));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; b++) {
            for (int c = 0; c < X->m; c++) hc[c] = FB_MUL(logit_hx[(int64_t)c * B + b], synd_x[(int64_t)c * B + b] ? -1.0f : 1.0f);
            for (int c = 0; c < Z->m; c++) hc[X->m + c] = FB_MUL(logit_hz[(int64_t)c * B + b], synd_z[(int64_t)c * B + b] ? -1.0f : 1.0f);
            for (int v = 0; v < n; v++)
                gnnd_vn(X, Z, G, hc, hc + X->m, v, h_vn + (b * n + v) * 3, out + (b * n + v) * 3);
        }
        free(hc);
    }
}

/* ---------------------------------------------------------------- OSD-0 ------------ */
/* OSD0_Decoder.call + find_mrb (bp_osd.py:8-77).  `rows`: the full-rank basis of the pcm (rank rows),
 * llr [n] reliabilities (small = likely in error), s [rank] reduced syndrome.  Columns are visited
 * in order of increasing llr (ties by index: the reference's tf.argsort leaves them unspecified);
 * for every row, in turn, the pivot is the first remaining one of that row, which is then eliminated
 * from all other rows (row operations on [pcm | s]).  e_hat is s-solution on the pivot columns. */
static void osd0_frame(const orc_rows_t *rows, int n, const float *llr, const uint8_t *s, uint8_t *e_hat,
                       int32_t *order, int32_t *inv, uint32_t *M) {
    const int R = rows->m, W = (n + 1 + 31) / 32;
    for (int v = 0; v < n; v++) order[v] = v;
    /* stable insertion-free sort: simple merge-free O(n log n) via qsort is not stable, so sort keys */
    for (int i = 1; i < n; i++) {            /* binary-insertion sort keeps ties in index order */
        int v = order[i];
        float key = llr[v];
        int lo = 0, hi = i;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (llr[order[mid]] <= key) lo = mid + 1; else hi = mid; }
        memmove(order + lo + 1, order + lo, sizeof(int32_t) * (size_t)(i - lo));
        order[lo] = v;
    }
    for (int j = 0; j < n; j++) inv[order[j]] = j;
    memset(M, 0, sizeof(uint32_t) * (size_t)R * W);
    for (int r = 0; r < R; r++) {
        for (int k = rows->ptr[r]; k < rows->ptr[r + 1]; k++) {
            int j = inv[rows->col[k]];
            M[(size_t)r * W + (j >> 5)] ^= 1u << (j & 31);
        }
        if (s[r]) M[(size_t)r * W + (n >> 5)] ^= 1u << (n & 31);
    }
    memset(e_hat, 0, (size_t)n);
    for (int r = 0; r < R; r++) {
        uint32_t *row = M + (size_t)r * W;
        int p = -1;
        for (int j = 0; j < n; j++) if ((row[j >> 5] >> (j & 31)) & 1u) { p = j; break; }
        if (p < 0) p = 0;                     /* tf.argmax of an all-zero row (cannot happen for a basis) */
        for (int i = 0; i < R; i++) {
            if (i == r) continue;
            uint32_t *ri = M + (size_t)i * W;
            if ((ri[p >> 5] >> (p & 31)) & 1u)
                for (int w = 0; w < W; w++) ri[w] ^= row[w];
        }
        order[n + r] = p;                     /* pivot of row r (permuted coordinates) */
    }
    for (int r = 0; r < R; r++) {
        int sol = (M[(size_t)r * W + (n >> 5)] >> (n & 31)) & 1u;
        e_hat[order[order[n + r]]] = (uint8_t)sol;      /* scatter to pivot, then undo the permutation */
    }
}

/* llr [B,n], s [rank,B], e_hat [B,n] */
void orc_osd0(const orc_rows_t *rows, int n, int64_t B, const float *llr, const uint8_t *s, uint8_t *e_hat) {
    const int R = rows->m, W = (n + 1 + 31) / 32;
#pragma omp parallel
    {
        int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + R + 1));
        int32_t *inv = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        uint32_t *M = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)R * W + 4);
        uint8_t *sr = (uint8_t *)malloc((size_t)R + 1);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; b++) {
            for (int r = 0; r < R; r++) sr[r] = s[(int64_t)r * B + b];
            osd0_frame(rows, n, llr + b * n, sr, e_hat + b * n, order, inv, M);
        }
        free(order); free(inv); free(M); free(sr);
    }
}

/* ---------------------------------------------------------------- pipelines -------- */
typedef struct {
    int32_t num_stages;          /* num_layers of the reference: 1 + number of GNN rounds     */
    const int32_t *num_iter;     /* [num_stages] BP iterations of decoders[i]                 */
    const float *factor;         /* [num_stages] normalization_factor of decoders[i]          */
    const int32_t *cn_type;      /* [num_stages]                                              */
    const orc_gnn_t *const *gnn; /* [num_stages-1] feedbacks[i]                               */
    float prior;                 /* log(3 (1 - p0) / p0), feedback_gnn.py:311-312              */
    int32_t fixed_weight;        /* > 0: Pauli(wt=True) errors of exactly this weight (300-301) */
    int32_t osd0;                /* 1: OSD-0 on the frames still mismatching after the last stage  */
                                 /*    (BP4_OSD_Model, bp_osd.py:80-197)                           */
    const orc_rows_t *basis_x;   /* hx[pivot_hx] */
    const int32_t *pivot_x;      /* [rank_x] rows of hx forming the basis */
    const orc_rows_t *basis_z;
    const int32_t *pivot_z;
    int32_t early_stop;          /* 1: opt-in early stop inside every BP stage (not the reference) */
    int32_t skip_inactive;       /* 0: every frame runs every round (reference behaviour);     */
                                 /* 1: stop a frame once its decision matches the syndrome     */
                                 /*    (result-identical, the scatter is masked: 339-340)      */
} orc_pipe_cfg_t;

/* flags bit0 = flagged (residual syndrome non-zero), bit1 = block error (ls_hat non-zero),
 * bits 2..7 = number of GNN rounds the frame was still active in. */
static uint8_t pipeline_frame(const orc_side_t *X, const orc_side_t *Z, const orc_rows_t *lx,
                              const orc_rows_t *lz, const orc_pipe_cfg_t *cfg,
                              const uint8_t *nx, const uint8_t *nz, float *fw, uint8_t *bw,
                              uint8_t *x_diff, uint8_t *z_diff) {
    const int n = X->n;
    float *mx = fw, *mz = mx + X->E, *pri = mz + Z->E, *L = pri + 3 * n, *xl = L + 3 * n,
          *zl = xl + Z->m, *wa = zl + X->m, *work = wa + 2 * n;
    uint8_t *sx = bw, *sz = sx + X->m, *xh = sz + Z->m, *zh = xh + n, *xh2 = zh + n,
            *zh2 = xh2 + n, *t = zh2 + n;
    const orc_rows_t rows_x = { Z->m, Z->cn_ptr, Z->cn_vn };   /* stage_one: _pcm_x_perp = hz */
    const orc_rows_t rows_z = { X->m, X->cn_ptr, X->cn_vn };   /*            _pcm_z_perp = hx */
    syndrome_frame(X, nz, sx);                 /* syndrome_x = hx . noise_z, feedback_gnn.py:308 */
    syndrome_frame(Z, nx, sz);                 /* syndrome_z = hz . noise_x                      */
    for (int v = 0; v < 3 * n; v++) pri[v] = cfg->prior;
    bp4_frame(X, Z, cfg->cn_type[0], cfg->num_iter[0], cfg->factor[0], pri, pri + n, pri + 2 * n,
              sx, sz, mx, mz, L, L + n, L + 2 * n, xh, zh, work);
    int active = 1, rounds = 0;
    for (int i = 1; i < cfg->num_stages; i++) {
        int mismatch = 0;                      /* feedback_gnn.py:324-330 */
        syndrome_frame(Z, xh, t);
        for (int c = 0; c < Z->m; c++) mismatch |= (t[c] != sz[c]);
        syndrome_frame(X, zh, t);
        for (int c = 0; c < X->m; c++) mismatch |= (t[c] != sx[c]);
        active = active && mismatch;
        if (!active && cfg->skip_inactive) break;
        if (active) rounds++;
        bp4_logits(n, &rows_x, &rows_z, L, L + n, L + 2 * n, xl, zl, wa, wa + n);
        /* feedbacks[i-1]((h_vn, logit_hz_perp, logit_hx_perp, ...)): z_logit pairs with hx */
        gnn_frame(X, Z, cfg->gnn[i - 1], L, L + n, L + 2 * n, zl, xl, sx, sz, pri, pri + n,
                  pri + 2 * n, work);
        bp4_frame(X, Z, cfg->cn_type[i], cfg->num_iter[i], cfg->factor[i], pri, pri + n,
                  pri + 2 * n, sx, sz, mx, mz, L, L + n, L + 2 * n, xh2, zh2, work);
        if (active) { memcpy(xh, xh2, (size_t)n); memcpy(zh, zh2, (size_t)n); }
    }
    if (cfg->osd0) {
        int mismatch = 0;
        syndrome_frame(Z, xh, t);
        for (int c = 0; c < Z->m; c++) mismatch |= (t[c] != sz[c]);
        syndrome_frame(X, zh, t);
        for (int c = 0; c < X->m; c++) mismatch |= (t[c] != sx[c]);
        if (mismatch) {
            /* bp_osd.py:135-144: osd_llrz = softplus(-llrx) - logsumexp(-llrz, -llry), osd_llrx likewise */
            float *lz_ = wa, *lx_ = wa + n;
            for (int v = 0; v < n; v++) {
                lz_[v] = FB_SUB(fb_m_softplusf(-L[v]), fb_m_logaddexpf(-L[2 * n + v], -L[n + v]));
                lx_[v] = FB_SUB(fb_m_softplusf(-L[2 * n + v]), fb_m_logaddexpf(-L[v], -L[n + v]));
            }
            const int Rx = cfg->basis_x->m, Rz = cfg->basis_z->m, Rm = Rx > Rz ? Rx : Rz;
            int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(2 * n + Rm + 1));
            uint32_t *M = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)Rm * ((n + 32) / 32) + 4);
            uint8_t *sr = (uint8_t *)malloc((size_t)Rm + 1);
            for (int r = 0; r < Rx; r++) sr[r] = sx[cfg->pivot_x[r]];
            osd0_frame(cfg->basis_x, n, lz_, sr, zh, order, order + n + Rm + 1, M);
            for (int r = 0; r < Rz; r++) sr[r] = sz[cfg->pivot_z[r]];
            osd0_frame(cfg->basis_z, n, lx_, sr, xh, order, order + n + Rm + 1, M);
            free(order); free(M); free(sr);
        }
    }
    for (int v = 0; v < n; v++) { xh[v] ^= nx[v]; zh[v] ^= nz[v]; }   /* x_diff, z_diff: 346-347 */
    if (x_diff) memcpy(x_diff, xh, (size_t)n);
    if (z_diff) memcpy(z_diff, zh, (size_t)n);
    int flagged = 0;
    syndrome_frame(Z, xh, t);
    for (int c = 0; c < Z->m; c++) flagged |= t[c];
    syndrome_frame(X, zh, t);
    for (int c = 0; c < X->m; c++) flagged |= t[c];
    /* any(hx_perp . x_diff) == any(hz . x_diff) or any(lz . x_diff)    */
    int blk = flagged | rows_any_parity(lz, xh) | rows_any_parity(lx, zh);
    return (uint8_t)((flagged ? 1 : 0) | (blk ? 2 : 0) | (rounds << 2));
}

static size_t pipe_float_work(const orc_side_t *X, const orc_side_t *Z, const orc_pipe_cfg_t *cfg) {
    size_t g = 0;
    for (int i = 0; i + 1 < cfg->num_stages; i++) {
        size_t w = gnn_work_floats(X, Z, cfg->gnn[i]);
        if (w > g) g = w;
    }
    int wd = max_cn_degree(X), wz = max_cn_degree(Z);
    if (wz > wd) wd = wz;
    if ((size_t)wd > g) g = (size_t)wd;
    return (size_t)X->E + (size_t)Z->E + 8 * (size_t)X->n + (size_t)X->m + (size_t)Z->m + g + 16;
}

/* Sandwich_BP_GNN_Evaluation_Model.call on frames [first_frame, first_frame + B).
 * thr: Pauli thresholds {px, px-py, (px+pz)-py} (float32).  If noise_x/noise_z are given
 * ([B,n] u8) they are used instead of sampling.  Outputs: flags [B]; counters[4] =
 * {frames, flagged, block errors, frames whose stage-0 decision missed the syndrome};
 * optional x_diff/z_diff [B,n]. */
void orc_pipeline(const orc_side_t *X, const orc_side_t *Z, const orc_rows_t *lx,
                  const orc_rows_t *lz, const orc_pipe_cfg_t *cfg, const float *thr,
                  uint64_t seed, uint64_t first_frame, int64_t B, const uint8_t *noise_x,
                  const uint8_t *noise_z, uint8_t *flags, int64_t *counters, uint8_t *x_diff,
                  uint8_t *z_diff) {
    const int n = X->n;
    int64_t c_flag = 0, c_blk = 0, c_s1 = 0;
    g_pipeline_early_stop = cfg->early_stop;
#pragma omp parallel reduction(+ : c_flag, c_blk, c_s1)
    {
        float *fw = (float *)malloc(sizeof(float) * pipe_float_work(X, Z, cfg));
        uint8_t *bw = (uint8_t *)malloc((size_t)(X->m + Z->m) * 2 + 8 * (size_t)n + 64);
        uint8_t *nx = (uint8_t *)malloc((size_t)n), *nz = (uint8_t *)malloc((size_t)n);
        uint16_t *idx = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; b++) {
            if (noise_x) { memcpy(nx, noise_x + b * n, (size_t)n); memcpy(nz, noise_z + b * n, (size_t)n); }
            else if (cfg->fixed_weight > 0) pauli_wt_frame(seed, first_frame + (uint64_t)b, n, cfg->fixed_weight, nx, nz, idx);
            else pauli_frame(seed, first_frame + (uint64_t)b, n, thr, nx, nz);
            uint8_t f = pipeline_frame(X, Z, lx, lz, cfg, nx, nz, fw, bw,
                                       x_diff ? x_diff + b * n : NULL, z_diff ? z_diff + b * n : NULL);
            if (flags) flags[b] = f;
            c_flag += f & 1; c_blk += (f >> 1) & 1; c_s1 += ((f >> 2) > 0);
        }
        free(fw); free(bw); free(nx); free(nz); free(idx);
    }
    if (counters) { counters[0] = B; counters[1] = c_flag; counters[2] = c_blk; counters[3] = c_s1; }
}

/* BP_BSC_Model.call (feedback_gnn.py:207-229) on one side: noise ~ Bernoulli(p), syndrome,
 * binary BP from the constant logit -log((1-p0)/p0), residual syndrome with the same pcm and
 * logical check with `logical` rows.  flags as in orc_pipeline (bits 0,1). */
void orc_bsc_pipeline(const orc_side_t *S, const orc_rows_t *logical, int cn_type, int num_iter,
                      float factor, float llr_const, float p, uint64_t seed, uint64_t first_frame,
                      int64_t B, const uint8_t *noise_in, uint8_t *flags, int64_t *counters,
                      const orc_rows_t *osd_basis, const int32_t *osd_pivot) {
    const int n = S->n, m = S->m;
    int wd = max_cn_degree(S);
    int64_t c_flag = 0, c_blk = 0;
#pragma omp parallel reduction(+ : c_flag, c_blk)
    {
        float *msg = (float *)malloc(sizeof(float) * (size_t)(S->E + 1));
        float *work = (float *)malloc(sizeof(float) * (size_t)wd);
        float *fl = (float *)malloc(sizeof(float) * 3 * (size_t)n);
        uint8_t *by = (uint8_t *)malloc(2 * (size_t)n + 2 * (size_t)m + 8);
        uint8_t *noise = by, *hard = by + n, *s = hard + n, *t = s + m;
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; b++) {
            if (noise_in) memcpy(noise, noise_in + b * n, (size_t)n);
            else bsc_frame(seed, first_frame + (uint64_t)b, n, p, noise);
            syndrome_frame(S, noise, s);
            for (int v = 0; v < n; v++) fl[v] = llr_const;
            bp2_frame(S, cn_type, num_iter, factor, fl, s, msg, fl + n, hard, work, fl + 2 * n);
            if (osd_basis) {                 /* BP2_OSD_Model, bp_osd.py:199-274 */
                int mismatch = 0;
                syndrome_frame(S, hard, t);
                for (int c = 0; c < m; c++) mismatch |= (t[c] != s[c]);
                if (mismatch) {
                    const int R = osd_basis->m;
                    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(2 * n + R + 1));
                    uint32_t *M = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)R * ((n + 32) / 32) + 4);
                    uint8_t *sr = (uint8_t *)malloc((size_t)R + 1);
                    for (int v = 0; v < n; v++) fl[2 * n + v] = -fl[n + v];       /* llr_hat = -decoder output */
                    for (int r = 0; r < R; r++) sr[r] = s[osd_pivot[r]];
                    osd0_frame(osd_basis, n, fl + 2 * n, sr, hard, order, order + n + R + 1, M);
                    free(order); free(M); free(sr);
                }
            }
            for (int v = 0; v < n; v++) hard[v] ^= noise[v];
            int flagged = 0;
            syndrome_frame(S, hard, t);
            for (int c = 0; c < m; c++) flagged |= t[c];
            int blk = logical ? rows_any_parity(logical, hard) : flagged;
            uint8_t f = (uint8_t)((flagged ? 1 : 0) | (blk ? 2 : 0));
            if (flags) flags[b] = f;
            c_flag += f & 1; c_blk += (f >> 1) & 1;
        }
        free(msg); free(work); free(fl); free(by);
    }
    if (counters) { counters[0] = B; counters[1] = c_flag; counters[2] = c_blk; counters[3] = 0; }
}

/* ---------------------------------------------------------------- GNN_BP4 ---------- */
/* GNN_BP4.call and UpdateCNEmbeddings / UpdateVNEmbeddings (gnn.py:71-751), configuration of
 * BASELINE configs[4]: 2-layer MLPs, no node/edge attributes.  As shipped, `call` unpacks five values
 * from cal_logit, which returns four (gnn.py:408 vs 314); the restatement drops the
 * fifth (`sum_llr`), which nothing else uses.  No weights and no recorded outputs of this layer exist
 * in the reference: parity of this component is UNPINNED (oracle-vs-CUDA only).
 *
 * MLP([x]) = W2^T act(W1^T x + b1) + b2.  Summation orders (the reference leaves them to the GEMM):
 * edge MLPs see x = [h_from (d), h_to (d)] and accumulate the receiver part (rows d..2d-1 of W1)
 * first, then the sender part (rows 0..d-1); everything else runs over ascending row index.  */
typedef struct {
    int32_t d, H, M, act, reduce, use_bias, num_iter;
    const float *Winv, *binv;                     /* _llr_inv_embed: [d,3], [3]                          */
    /* update_h_cn: msg_mlp_x, msg_mlp_z ([2d,H],[H],[H,M],[M]); embed_mlp_x, embed_mlp_z ([M+d+1,H],[H],[H,d],[d]) */
    const float *cmx[4], *cmz[4], *cex[4], *cez[4];
    /* update_h_vn: msg_mlp_x, msg_mlp_z; embed_mlp ([2M+d,H],[H],[H,d],[d]) */
    const float *vmx[4], *vmz[4], *ve[4];
} orc_gbp_t;

static void gbp_edge_mlp(const orc_gbp_t *G, const float *const w[4], const float *from, const float *base,
                         float *hid, float *msg) {
    const int d = G->d, H = G->H, M = G->M;
    for (int j = 0; j < H; j++) {
        float a = base[j];
        for (int k = 0; k < d; k++) a = FB_FMA(from[k], w[0][k * H + j], a);
        if (w[1]) a = FB_ADD(a, w[1][j]);
        hid[j] = gnn_act(G->act, a);
    }
    for (int i = 0; i < M; i++) {
        float a = 0.0f;
        for (int j = 0; j < H; j++) a = FB_FMA(hid[j], w[2][j * M + i], a);
        if (w[3]) a = FB_ADD(a, w[3][i]);
        msg[i] = a;
    }
}

static void gbp_base(const orc_gbp_t *G, const float *const w[4], const float *to, float *base) {
    const int d = G->d, H = G->H;
    for (int j = 0; j < H; j++) {
        float a = 0.0f;
        for (int k = 0; k < d; k++) a = FB_FMA(to[k], w[0][(d + k) * H + j], a);
        base[j] = a;
    }
}

static void gbp_node_mlp(const orc_gbp_t *G, const float *const w[4], const float *in, int K, float *hid, float *out) {
    const int d = G->d, H = G->H;
    for (int j = 0; j < H; j++) {
        float a = 0.0f;
        for (int k = 0; k < K; k++) a = FB_FMA(in[k], w[0][k * H + j], a);
        if (w[1]) a = FB_ADD(a, w[1][j]);
        hid[j] = gnn_act(G->act, a);
    }
    for (int i = 0; i < d; i++) {
        float a = 0.0f;
        for (int j = 0; j < H; j++) a = FB_FMA(hid[j], w[2][j * d + i], a);
        if (w[3]) a = FB_ADD(a, w[3][i]);
        out[i] = a;
    }
}

static void gbp_reduce(const orc_gbp_t *G, float *red, const float *msg, int first, int M) {
    for (int i = 0; i < M; i++) {
        if (first) red[i] = msg[i];
        else if (G->reduce == 2) red[i] = (msg[i] > red[i]) ? msg[i] : red[i];
        else red[i] = (msg[i] < red[i]) ? msg[i] : red[i];
    }
}

/* reduce_op "mean" / "sum": both layers of the edge MLP are applied in factored form.  The first layer is
 * linear in [h_from, h_to], so its two halves are formed per NODE (sender half Pf = W1[:d]^T h_from,
 * receiver half base = W1[d:]^T h_to) and only added per edge; the output layer is linear, so it is applied
 * once to the signed sum of the hidden activations over the receiver's edges:
 *     sum_e s_e (W2^T t_e + b2) = W2^T (sum_e s_e t_e) + (sum_e s_e) b2        (mean: divided by deg).
 * Same value as the reference's per-edge Dense layers up to float32 re-association (np_oracle.py keeps the
 * literal form; tests/test_gnn_bp4.py bounds the difference).  "max" / "min" use the per-edge form above. */
static void gbp_from(const orc_gbp_t *G, const float *const w[4], const float *from, float *pf) {
    const int d = G->d, H = G->H;
    for (int j = 0; j < H; j++) {
        float a = 0.0f;
        for (int k = 0; k < d; k++) a = FB_FMA(from[k], w[0][k * H + j], a);
        pf[j] = a;
    }
}

/* hs[j] (+)= s * act(pf[j] + base[j] + b1[j]) */
static void gbp_hidden_acc(const orc_gbp_t *G, const float *const w[4], const float *pf, const float *base, int negate,
                           int first, float *hs) {
    for (int j = 0; j < G->H; j++) {
        float a = FB_ADD(pf[j], base[j]);
        if (w[1]) a = FB_ADD(a, w[1][j]);
        float t = gnn_act(G->act, a);
        if (negate) t = -t;
        hs[j] = first ? t : FB_ADD(hs[j], t);
    }
}

/* red[i] = (sum_j hs[j] W2[j][i] + ssum b2[i]) (/ deg for "mean") */
static void gbp_out_layer(const orc_gbp_t *G, const float *const w[4], const float *hs, float ssum, int deg, float *red) {
    const int H = G->H, M = G->M;
    for (int i = 0; i < M; i++) {
        float a = 0.0f;
        for (int j = 0; j < H; j++) a = FB_FMA(hs[j], w[2][j * M + i], a);
        if (w[3]) a = FB_FMA(ssum, w[3][i], a);
        if (G->reduce == 0 && deg > 0) a = FB_DIV(a, (float)deg);
        red[i] = a;
    }
}

/* UpdateCNEmbeddings.call for one side (gnn.py:574-610) */
static void gbp_cn_update(const orc_gbp_t *G, const orc_side_t *S, const float *const wm[4], const float *const we[4],
                          const float *h_vn, float *h_cn, const float *logit, float *work) {
    const int d = G->d, H = G->H, M = G->M;
    float *base = work, *hid = base + H, *msg = hid + H, *red = msg + M, *in = red + M, *out = in + M + d + 1,
          *pf = out + d, *hs = pf + H;
    for (int c = 0; c < S->m; c++) {
        gbp_base(G, wm, h_cn + c * d, base);
        const int k0 = S->cn_ptr[c], k1 = S->cn_ptr[c + 1];
        for (int i = 0; i < M; i++) red[i] = 0.0f;
        if (G->reduce <= 1) {
            for (int j = 0; j < H; j++) hs[j] = 0.0f;
            for (int k = k0; k < k1; k++) {
                gbp_from(G, wm, h_vn + S->cn_vn[k] * d, pf);
                gbp_hidden_acc(G, wm, pf, base, 0, k == k0, hs);
            }
            gbp_out_layer(G, wm, hs, (float)(k1 - k0), k1 - k0, red);
        } else {
            for (int k = k0; k < k1; k++) {
                gbp_edge_mlp(G, wm, h_vn + S->cn_vn[k] * d, base, hid, msg);
                gbp_reduce(G, red, msg, k == k0, M);
            }
        }
        for (int i = 0; i < M; i++) in[i] = red[i];
        for (int i = 0; i < d; i++) in[M + i] = h_cn[c * d + i];
        in[M + d] = logit[c];
        gbp_node_mlp(G, we, in, M + d + 1, hid, out);
        for (int i = 0; i < d; i++) h_cn[c * d + i] = out[i];
    }
}

/* UpdateVNEmbeddings.call (gnn.py:716-750) */
static void gbp_vn_update(const orc_gbp_t *G, const orc_side_t *X, const orc_side_t *Z, const float *hcx,
                          const float *hcz, float *h_vn, const uint8_t *sx, const uint8_t *sz, float *work) {
    const int d = G->d, H = G->H, M = G->M;
    float *base = work, *hid = base + H, *msg = hid + H, *in = msg + M, *out = in + 2 * M + d, *pf = out + d, *hs = pf + H;
    for (int v = 0; v < X->n; v++) {
        for (int side = 0; side < 2; side++) {
            const orc_side_t *S = side ? Z : X;
            const float *const *wm = side ? G->vmz : G->vmx;
            const float *hc = side ? hcz : hcx;
            const uint8_t *sy = side ? sz : sx;
            float *red = in + side * M;
            gbp_base(G, wm, h_vn + v * d, base);
            const int e0 = S->vn_ptr[v], e1 = S->vn_ptr[v + 1];
            for (int i = 0; i < M; i++) red[i] = 0.0f;
            if (G->reduce <= 1) {
                int ssum = 0;
                for (int j = 0; j < H; j++) hs[j] = 0.0f;
                for (int e = e0; e < e1; e++) {
                    const int c = S->vn_cn[e];
                    gbp_from(G, wm, hc + c * d, pf);
                    gbp_hidden_acc(G, wm, pf, base, sy[c] != 0, e == e0, hs);    /* x (1 - 2 s), gnn.py:733-737 */
                    ssum += sy[c] ? -1 : 1;
                }
                gbp_out_layer(G, wm, hs, (float)ssum, e1 - e0, red);
                continue;
            }
            for (int e = e0; e < e1; e++) {
                const int c = S->vn_cn[e];
                gbp_edge_mlp(G, wm, hc + c * d, base, hid, msg);
                if (sy[c]) for (int i = 0; i < M; i++) msg[i] = -msg[i];          /* x (1 - 2 s), gnn.py:733-737 */
                gbp_reduce(G, red, msg, e == e0, M);
            }
        }
        for (int i = 0; i < d; i++) in[2 * M + i] = h_vn[v * d + i];
        gbp_node_mlp(G, G->ve, in, 2 * M + d, hid, out);
        for (int i = 0; i < d; i++) h_vn[v * d + i] = out[i];
    }
}

static float gbp_soft_row(const orc_rows_t *R, int r, const float *l) {
    float sgn = 1.0f, T = 0.0f;
    for (int k = R->ptr[r]; k < R->ptr[r + 1]; k++) {
        float m = l[R->col[k]];
        if (m < 0.0f) sgn = -sgn;
        T = FB_ADD(T, fb_m_phi2f(fabsf(m)));
    }
    return FB_MUL(sgn, fb_m_phi2f(T));
}

/* One frame of GNN_BP4.call.  sx [m_x], sz [m_z]; x_logit [num_iter][m_z + k_z], z_logit [num_iter][m_x + k_x]
 * (x_perp_logit = [hz_logit; lz_logit], z_perp_logit = [hx_logit; lx_logit], gnn.py:311-313); xh, zh [n]. */
static void gbp_frame(const orc_gbp_t *G, const orc_side_t *X, const orc_side_t *Z, const orc_rows_t *lx,
                      const orc_rows_t *lz, const uint8_t *sx, const uint8_t *sz, float *x_logit, float *z_logit,
                      uint8_t *xh, uint8_t *zh, float *fw) {
    const int n = X->n, d = G->d;
    const orc_rows_t hxr = { X->m, X->cn_ptr, X->cn_vn }, hzr = { Z->m, Z->cn_ptr, Z->cn_vn };
    float *h_vn = fw, *hcx = h_vn + n * d, *hcz = hcx + X->m * d, *lgx = hcz + Z->m * d, *lgz = lgx + X->m,
          *llr = lgz + Z->m, *lxp = llr + 3 * n, *lzp = lxp + n, *work = lzp + n;
    for (int i = 0; i < n * d; i++) h_vn[i] = 1.0f;
    for (int i = 0; i < (X->m + Z->m) * d; i++) hcx[i] = 0.0f;
    for (int c = 0; c < X->m + Z->m; c++) lgx[c] = 0.0f;
    gbp_cn_update(G, X, G->cmx, G->cex, h_vn, hcx, lgx, work);
    gbp_cn_update(G, Z, G->cmz, G->cez, h_vn, hcz, lgz, work);
    for (int it = 0; it < G->num_iter; it++) {
        gbp_vn_update(G, X, Z, hcx, hcz, h_vn, sx, sz, work);
        for (int v = 0; v < n; v++) {                       /* embed_to_llr + cal_logit, gnn.py:281-314 */
            for (int c = 0; c < 3; c++) {
                float a = 0.0f;
                for (int k = 0; k < d; k++) a = FB_FMA(h_vn[v * d + k], G->Winv[k * 3 + c], a);
                if (G->binv) a = FB_ADD(a, G->binv[c]);
                llr[c * n + v] = a;
            }
            const float Lx = llr[v], Ly = llr[n + v], Lz = llr[2 * n + v];
            lzp[v] = FB_SUB(fb_m_softplusf(-Lx), fb_m_logaddexpf(-Lz, -Ly));
            lxp[v] = FB_SUB(fb_m_softplusf(-Lz), fb_m_logaddexpf(-Lx, -Ly));
        }
        float *xo = x_logit + (size_t)it * (Z->m + lz->m), *zo = z_logit + (size_t)it * (X->m + lx->m);
        for (int r = 0; r < Z->m; r++) xo[r] = lgz[r] = gbp_soft_row(&hzr, r, lxp);
        for (int r = 0; r < lz->m; r++) xo[Z->m + r] = gbp_soft_row(lz, r, lxp);
        for (int r = 0; r < X->m; r++) zo[r] = lgx[r] = gbp_soft_row(&hxr, r, lzp);
        for (int r = 0; r < lx->m; r++) zo[X->m + r] = gbp_soft_row(lx, r, lzp);
        if (it == G->num_iter - 1) break;
        for (int c = 0; c < X->m; c++) if (sx[c]) lgx[c] = -lgx[c];            /* hx_logit * (1 - 2 s) */
        for (int c = 0; c < Z->m; c++) if (sz[c]) lgz[c] = -lgz[c];
        gbp_cn_update(G, X, G->cmx, G->cex, h_vn, hcx, lgx, work);
        gbp_cn_update(G, Z, G->cmz, G->cez, h_vn, hcz, lgz, work);
    }
    for (int v = 0; v < n; v++) {                           /* make_hard_decision, gnn.py:358-366 */
        int dd = 0;
        float best = 0.0f;
        if (llr[v] < best) { best = llr[v]; dd = 1; }
        if (llr[2 * n + v] < best) { best = llr[2 * n + v]; dd = 2; }
        if (llr[n + v] < best) { best = llr[n + v]; dd = 3; }
        xh[v] = (uint8_t)(dd & 1);
        zh[v] = (uint8_t)(dd >> 1);
    }
}

/* synd_x [B,m_x], synd_z [B,m_z] (batch first, gnn.py:385-386); x_logit [num_iter][m_z+k_z][B],
 * z_logit [num_iter][m_x+k_x][B]; x_hat, z_hat [n][B] */
void orc_gnn_bp4(const orc_gbp_t *G, const orc_side_t *X, const orc_side_t *Z, const orc_rows_t *lx,
                 const orc_rows_t *lz, int64_t B, const uint8_t *synd_x, const uint8_t *synd_z, float *x_logit,
                 float *z_logit, uint8_t *x_hat, uint8_t *z_hat) {
    const int n = X->n, d = G->d, rx = Z->m + lz->m, rz = X->m + lx->m;
#pragma omp parallel
    {
        const size_t fwn = (size_t)(n + X->m + Z->m) * d + X->m + Z->m + 5 * (size_t)n + 6 * (size_t)G->H +
                           6 * (size_t)G->M + 4 * (size_t)d + 64;
        float *fw = (float *)malloc(sizeof(float) * fwn);
        float *xl = (float *)malloc(sizeof(float) * (size_t)G->num_iter * rx + 4);
        float *zl = (float *)malloc(sizeof(float) * (size_t)G->num_iter * rz + 4);
        uint8_t *xh = (uint8_t *)malloc(2 * (size_t)n);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; b++) {
            gbp_frame(G, X, Z, lx, lz, synd_x + b * X->m, synd_z + b * Z->m, xl, zl, xh, xh + n, fw);
            for (int it = 0; it < G->num_iter; it++) {
                for (int r = 0; r < rx; r++) x_logit[((int64_t)it * rx + r) * B + b] = xl[(size_t)it * rx + r];
                for (int r = 0; r < rz; r++) z_logit[((int64_t)it * rz + r) * B + b] = zl[(size_t)it * rz + r];
            }
            for (int v = 0; v < n; v++) { x_hat[(int64_t)v * B + b] = xh[v]; z_hat[(int64_t)v * B + b] = xh[n + v]; }
        }
        free(fw); free(xl); free(zl); free(xh);
    }
}

/* ---------------------------------------------------------------- math probes ------ */
#define ORC_VEC(name, fn) \
    void name(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = fn(x[i]); }
ORC_VEC(orc_expf, fb_expf)
ORC_VEC(orc_logf, fb_logf)
ORC_VEC(orc_log1pf_pos, fb_log1pf_pos)
ORC_VEC(orc_softplusf, fb_softplusf)
ORC_VEC(orc_phi4f, fb_phi4f)
ORC_VEC(orc_phi2f, fb_phi2f)
ORC_VEC(orc_tanhf, fb_tanhf)
ORC_VEC(orc_atanhf, fb_atanhf)
/* in the arithmetic selected by orc_set_math */
ORC_VEC(orc_m_softplusf, fb_m_softplusf)
ORC_VEC(orc_m_phi4f, fb_m_phi4f)
ORC_VEC(orc_m_phi2f, fb_m_phi2f)
ORC_VEC(orc_sfu_expf, fb_sfu_expf)
ORC_VEC(orc_sfu_logf, fb_sfu_logf)
ORC_VEC(orc_m_tanhf, fb_m_tanhf)
ORC_VEC(orc_m_atanhf, fb_m_atanhf)
void orc_logaddexpf(const float *a, const float *b, float *y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = fb_m_logaddexpf(a[i], b[i]);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* arithmetic of every oracle function from now on: 0 = exact (software libm, default), 1 = SFU (MUFU tables;
 * needs orc_set_sfu_tables first).  See fb_math.h. */
int orc_set_math(int mode) {
    if (mode != 0 && mode != 1) return -1;
    if (mode == 1 && (!fb_sfu_ex2_tab || !fb_sfu_lg2_tab || !fb_sfu_lg2b_tab || !fb_sfu_rcp_tab)) return -2;
    fb_math_mode = mode;
    return 0;
}
int orc_get_math(void) { return fb_math_mode; }
void orc_set_sfu_tables(const float *ex2_tab, const float *lg2_tab, const float *lg2b_tab, const float *rcp_tab) {
    fb_sfu_set_tables(ex2_tab, lg2_tab, lg2b_tab, rcp_tab);
}

void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}


/* ---------------------------------------------------------------- tensor-core step model ---------------- */
/* Replays a dump of tools/micro/umma_probe.cu (A [T,128,8], B [T,8,16], Din / Dout [T,128,16]) through fb_umma8 and
 * returns the number of outputs whose bits differ from the hardware's. */
int64_t orc_umma8_check(const float *A, const float *B, const float *Din, const float *Dout, int32_t trials) {
    int64_t bad = 0;
    for (int tr = 0; tr < trials; tr++)
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 16; n++) {
                const float got = fb_umma8(Din[((int64_t)tr * 128 + m) * 16 + n], 1, A + ((int64_t)tr * 128 + m) * 8,
                                           B + (int64_t)tr * 128 + n, 16);
                uint32_t g, w;
                memcpy(&g, &got, 4); memcpy(&w, &Dout[((int64_t)tr * 128 + m) * 16 + n], 4);
                bad += g != w;
            }
    return bad;
}
