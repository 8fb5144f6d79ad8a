// fbgnn_train.cuh -- gradient of the second training stage (feedback GNN -> BP4 with per-iteration soft
// syndromes -> multi-loss BCE), i.e. what tf.GradientTape computes around Second_Stage_GNN_BP_Model.call
// (sionna/fec/ldpc/feedback_gnn.py:395-460; training loop in examples/Feedback_GNN.ipynb cells 2, 8).
//
//   loss = sum_{i = loss_from}^{num_iter-1} [ bce(1 - s_z, x_logit_{i+1}) + bce(1 - s_x, z_logit_{i+1}) ]
//
// with x_logit_k / z_logit_k the soft syndromes of the marginals after k check-node updates (llr_hat[2k],
// llr_hat[2k+1] of decoding_q.py:743-746,779-780), bce = tf.keras BinaryCrossentropy(from_logits=True) (mean
// over all entries), and the priors of the decoder the output of the feedback GNN.
//
// k_bp4_grad     one CTA per frame: forward BP4 (boxplus-phi) keeping the check-to-variable messages of every
//                iteration in an HBM trace, then the reverse sweep; returns d loss / d priors [B,3,n] and the
//                frame's share of the loss.
// k_gnn_bwd_rows one thread per (frame, variable node): recomputes the GNN forward (factored form) and writes
//                the per-row factors of every weight gradient.
// k_atb_partial / k_atb_reduce   dW = A^T B over all rows, deterministic two-stage reduction.
//
// Derivatives follow TensorFlow's: sign() and the comparison ops carry no gradient, clip_by_value passes the
// gradient inside the clip range only, phi'(x) = sigmoid(x) - e^x / (e^x - 1) = -1 / sinh(x).
#ifndef FBGNN_TRAIN_CUH
#define FBGNN_TRAIN_CUH

#include "fbgnn_kernels.cuh"

namespace fbgnn {
namespace train {

struct Bp4GradArgs {
    SideDev X, Z;
    int num_iter, loss_from;
    float factor;
    int64_t B;
    View3<const float> llr;             // priors (b, k, v), k = x, y, z
    View2<const uint8_t> sx, sz;        // (c, b)
    float *trace;                       // [B][num_iter][E_x + E_z]: c2v messages after 1..num_iter updates
    float *dprior;                      // [B][3][n] out (nullptr: loss only)
    double *loss;                       // [B] out: this frame's share of the loss
    float wx, wz;                       // 1 / (B m_z), 1 / (B m_x): the means of the two BCE terms
};

__device__ __forceinline__ float dphi(float x) {             // derivative of phi (with its clip), see header
    if (!(x >= FB_PHI_CLIP_LO && x <= FB_PHI_CLIP_HI)) return 0.0f;
    return -1.0f / sinhf(x);
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// marginals of one variable node from the current c2v messages
__device__ __forceinline__ void marginals(const SideDev &X, const SideDev &Z, const float *mx, const float *mz,
                                          const float *pri, int n, int v, float &lx, float &ly, float &lz) {
    float Sx = 0.0f, Sz = 0.0f;
    for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) Sx = FB_ADD(Sx, mx[e]);
    for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) Sz = FB_ADD(Sz, mz[e]);
    ly = FB_ADD(FB_ADD(Sz, Sx), pri[n + v]);
    lx = FB_ADD(Sz, pri[v]);
    lz = FB_ADD(Sx, pri[2 * n + v]);
}

// smem (floats): mx[Ex] mz[Ez] | vx[Ex] vz[Ez] | gx[Ex] gz[Ez] | pri[3n] | L[3n] | dL[3n] | dP[3n] | lp[2n] | dT[m]
__global__ void __launch_bounds__(256) k_bp4_grad(const Bp4GradArgs a) {
    typedef MathExact MATH;
    extern __shared__ float gsm[];
    const SideDev &X = a.X, &Z = a.Z;
    const int n = X.n, Ex = X.E, Ez = Z.E, E = Ex + Ez, mt = X.m + Z.m, T = blockDim.x, tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    float *mx = gsm, *mz = mx + Ex, *vx = mz + Ez, *vz = vx + Ex, *gx = vz + Ez, *gz = gx + Ex;
    float *pri = gz + Ez, *L = pri + 3 * n, *dL = L + 3 * n, *dP = dL + 3 * n, *lp = dP + 3 * n, *dT = lp + 2 * n;
    float *trace = a.trace + b * (int64_t)a.num_iter * E;
    for (int i = tid; i < 3 * n; i += T) { pri[i] = a.llr(b, i / n, i % n); dP[i] = 0.0f; }
    for (int e = tid; e < E; e += T) { mx[e] = 0.0f; gx[e] = 0.0f; }
    __syncthreads();

    // ---------------- forward: num_iter iterations, c2v messages of every iteration kept
    for (int it = 0; it < a.num_iter; it++) {
        for (int v = tid; v < n; v += T) {
            float lx, ly, lz;
            marginals(X, Z, mx, mz, pri, n, v, lx, ly, lz);
            const float spx = MATH::softplus(-lx), spz = MATH::softplus(-lz);
            for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) {
                const float m = mx[e];
                vx[e] = FB_SUB(spx, MATH::logaddexp(-FB_SUB(lz, m), -FB_SUB(ly, m)));
            }
            for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) {
                const float m = mz[e];
                vz[e] = FB_SUB(spz, MATH::logaddexp(-FB_SUB(lx, m), -FB_SUB(ly, m)));
            }
        }
        __syncthreads();
        for (int c = tid; c < mt; c += T) {
            const bool isx = c < X.m;
            const SideDev &S = isx ? X : Z;
            const int cc = isx ? c : c - X.m;
            const float *vin = isx ? vx : vz;
            float *mout = isx ? mx : mz;
            const int k0 = S.cn_ptr[cc], k1 = S.cn_ptr[cc + 1];
            int par = isx ? a.sx(cc, b) : a.sz(cc, b);
            float Ts = 0.0f;
            for (int k = k0; k < k1; k++) {
                const float m = vin[S.cn_edge[k]];
                par ^= (m < 0.0f);
                Ts = FB_ADD(Ts, MATH::phi4(fabsf(m)));
            }
            for (int k = k0; k < k1; k++) {
                const int e = S.cn_edge[k];
                const float m = vin[e];
                float o = MATH::phi4(FB_SUB(Ts, MATH::phi4(fabsf(m))));
                if (par ^ (int)(m < 0.0f)) o = -o;
                mout[e] = FB_MUL(o, a.factor);
            }
        }
        __syncthreads();
        for (int e = tid; e < E; e += T) trace[(int64_t)it * E + e] = mx[e];
    }

    // ---------------- reverse sweep over the message states k = num_iter .. 0
    double loss = 0.0;
    for (int k = a.num_iter; k >= 0; k--) {
        __syncthreads();
        for (int e = tid; e < E; e += T) mx[e] = k > 0 ? trace[(int64_t)(k - 1) * E + e] : 0.0f;
        __syncthreads();
        const bool has_loss = k >= a.loss_from + 1;
        for (int v = tid; v < n; v += T) {
            float lx, ly, lz;
            marginals(X, Z, mx, mz, pri, n, v, lx, ly, lz);
            L[v] = lx; L[n + v] = ly; L[2 * n + v] = lz;
            dL[v] = 0.0f; dL[n + v] = 0.0f; dL[2 * n + v] = 0.0f;
            if (has_loss) {
                lp[v] = FB_SUB(MATH::softplus(-lz), MATH::logaddexp(-lx, -ly));          // llr_x'
                lp[n + v] = FB_SUB(MATH::softplus(-lx), MATH::logaddexp(-lz, -ly));      // llr_z'
            }
        }
        __syncthreads();
        if (has_loss) {
            // soft syndromes: x_logit over rows of hz (from llr_x'), z_logit over rows of hx (from llr_z')
            for (int c = tid; c < mt; c += T) {
                const bool isx = c < X.m;                       // row of hx -> z_logit, label 1 - s_x
                const SideDev &S = isx ? X : Z;
                const int cc = isx ? c : c - X.m;
                const float *l = isx ? lp + n : lp;
                float Ts = 0.0f;
                int neg = 0;
                for (int q = S.cn_ptr[cc]; q < S.cn_ptr[cc + 1]; q++) {
                    const float m = l[S.cn_vn[q]];
                    neg ^= (m < 0.0f);
                    Ts = FB_ADD(Ts, MATH::phi4(fabsf(m)));
                }
                float logit = MATH::phi4(Ts);
                if (neg) logit = -logit;
                const float z = 1.0f - (float)(isx ? a.sx(cc, b) : a.sz(cc, b));
                const float w = isx ? a.wz : a.wx;
                loss += (double)w * ((double)fmaxf(logit, 0.0f) - (double)logit * z + (double)log1pf(__expf(-fabsf(logit))));
                const float dlogit = w * (sigmoidf(logit) - z);
                dT[c] = dlogit * (neg ? -1.0f : 1.0f) * dphi(Ts);
            }
            __syncthreads();
            for (int v = tid; v < n; v += T) {
                float sx_ = 0.0f, sz_ = 0.0f;                   // sum of dT over the rows containing v
                for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) sx_ += dT[X.m + Z.vn_cn[e]];     // hz rows use llr_x'
                for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) sz_ += dT[X.vn_cn[e]];           // hx rows use llr_z'
                const float lxp = lp[v], lzp = lp[n + v];
                const float dlxp = sx_ * dphi(fabsf(lxp)) * (lxp < 0.0f ? -1.0f : 1.0f);
                const float dlzp = sz_ * dphi(fabsf(lzp)) * (lzp < 0.0f ? -1.0f : 1.0f);
                const float lx = L[v], ly = L[n + v], lz = L[2 * n + v];
                // llr_x' = softplus(-lz) - logaddexp(-lx, -ly);  llr_z' = softplus(-lx) - logaddexp(-lz, -ly)
                const float wxy = sigmoidf(ly - lx);            // weight of the -lx term in logaddexp(-lx, -ly)
                const float wzy = sigmoidf(ly - lz);
                dL[v] += dlxp * wxy - dlzp * sigmoidf(-lx);
                dL[n + v] += dlxp * (1.0f - wxy) + dlzp * (1.0f - wzy);
                dL[2 * n + v] += -dlxp * sigmoidf(-lz) + dlzp * wzy;
            }
            __syncthreads();
        }
        if (k < a.num_iter) {
            // v2c messages of state k (they produced the c2v messages of state k+1, whose gradient is gx / gz)
            for (int v = tid; v < n; v += T) {
                const float lx = L[v], ly = L[n + v], lz = L[2 * n + v];
                const float spx = MATH::softplus(-lx), spz = MATH::softplus(-lz);
                for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) {
                    const float m = mx[e];
                    vx[e] = FB_SUB(spx, MATH::logaddexp(-FB_SUB(lz, m), -FB_SUB(ly, m)));
                }
                for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) {
                    const float m = mz[e];
                    vz[e] = FB_SUB(spz, MATH::logaddexp(-FB_SUB(lx, m), -FB_SUB(ly, m)));
                }
            }
            __syncthreads();
            // check-node update, reverse: gradient w.r.t. its inputs, written over vx / vz
            for (int c = tid; c < mt; c += T) {
                const bool isx = c < X.m;
                const SideDev &S = isx ? X : Z;
                const int cc = isx ? c : c - X.m;
                float *vin = isx ? vx : vz;
                const float *g = isx ? gx : gz;
                const int k0 = S.cn_ptr[cc], k1 = S.cn_ptr[cc + 1];
                int par = isx ? a.sx(cc, b) : a.sz(cc, b);
                float Ts = 0.0f;
                for (int q = k0; q < k1; q++) {
                    const float m = vin[S.cn_edge[q]];
                    par ^= (m < 0.0f);
                    Ts = FB_ADD(Ts, MATH::phi4(fabsf(m)));
                }
                float Q = 0.0f;
                for (int q = k0; q < k1; q++) {
                    const int e = S.cn_edge[q];
                    const float m = vin[e];
                    const float sg = (par ^ (int)(m < 0.0f)) ? -1.0f : 1.0f;
                    Q += g[e] * sg * a.factor * dphi(FB_SUB(Ts, MATH::phi4(fabsf(m))));
                }
                for (int q = k0; q < k1; q++) {
                    const int e = S.cn_edge[q];
                    const float m = vin[e];
                    const float am = fabsf(m);
                    const float sg = (par ^ (int)(m < 0.0f)) ? -1.0f : 1.0f;
                    const float qe = g[e] * sg * a.factor * dphi(FB_SUB(Ts, MATH::phi4(am)));
                    vin[e] = (Q - qe) * dphi(am) * (m < 0.0f ? -1.0f : 1.0f);
                }
            }
            __syncthreads();
            // variable-node update, reverse: into dL; the explicit -m dependence goes to gx / gz
            for (int v = tid; v < n; v += T) {
                const float lx = L[v], ly = L[n + v], lz = L[2 * n + v];
                float ax = 0.0f, ay = 0.0f, az = 0.0f;
                const float nsx = -sigmoidf(-lx), nsz = -sigmoidf(-lz);
                for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) {
                    // vx = softplus(-lx) - logaddexp(-(lz - m), -(ly - m)); weights of the two terms do not depend on m
                    const float h = vx[e], w1 = sigmoidf(ly - lz);
                    ax += h * nsx; az += h * w1; ay += h * (1.0f - w1);
                    gx[e] = -h;
                }
                for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) {
                    const float h = vz[e], w1 = sigmoidf(ly - lx);
                    az += h * nsz; ax += h * w1; ay += h * (1.0f - w1);
                    gz[e] = -h;
                }
                dL[v] += ax; dL[n + v] += ay; dL[2 * n + v] += az;
            }
        } else {
            for (int e = tid; e < E; e += T) gx[e] = 0.0f;
        }
        __syncthreads();
        // marginals -> priors and messages of state k
        for (int v = tid; v < n; v += T) {
            const float dx = dL[v], dy = dL[n + v], dz = dL[2 * n + v];
            dP[v] += dx; dP[n + v] += dy; dP[2 * n + v] += dz;
            const float dSx = dz + dy, dSz = dx + dy;
            for (int e = X.vn_ptr[v]; e < X.vn_ptr[v + 1]; e++) gx[e] += dSx;
            for (int e = Z.vn_ptr[v]; e < Z.vn_ptr[v + 1]; e++) gz[e] += dSz;
        }
    }
    __syncthreads();
    if (a.dprior) for (int i = tid; i < 3 * n; i += T) a.dprior[b * 3 * n + i] = dP[i];
    // the frame's loss: block reduction in double
    __shared__ double red[256];
    red[tid] = loss;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    if (tid == 0) a.loss[b] = red[0];
}

// ------------------------------------------------------------------ feedback GNN, reverse
// Row factors of the weight gradients ("1" columns make the bias gradients part of the same products):
//   hid1 [R][H+1]   (tanh of the VN MLP's hidden layer | 1)       x  dout  [R][3]     -> [W0; b0]
//   in1  [R][2M+4]  (m_x, m_z, Lx, Ly, Lz | 1)                    x  dpre3 [R][H]     -> [W3; b3]
//   hs1  [2][R][H+1](sum of edge activations / deg | bias factor) x  dm    [2][R][M]  -> [W2s; b2s]
//   ft1  [Es rows][5] (h_cn, Lx, Ly, Lz | 1)                      x  dpre1 [Es rows][H] -> [W1s; b1s]
struct GnnBwdArgs {
    SideDev X, Z;
    const float *weights;               // packed, GnnLayout
    int reduce;                         // 0 mean, 1 sum
    int64_t B;
    View3<const float> h_vn;
    View2<const float> logit_hx, logit_hz;
    View2<const uint8_t> sx, sz;
    const float *dout;                  // [B][3][n]
    float *hid1, *dpre3, *in1, *dout_r, *hs1, *dm, *ftx, *dpx, *ftz, *dpz;
};

template <int H, int M>
__global__ void __launch_bounds__(128) k_gnn_bwd_rows(const GnnBwdArgs a) {
    typedef GnnLayout<H, M> Lay;
    extern __shared__ float wsm[];
    for (int i = threadIdx.x; i < Lay::total; i += blockDim.x) wsm[i] = a.weights[i];
    __syncthreads();
    const float *w = wsm;
    const int n = a.X.n;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.B * n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = r / n;
        const int v = (int)(r - b * n);
        const float f[3] = {a.h_vn(b, v, 0), a.h_vn(b, v, 1), a.h_vn(b, v, 2)};
        float in[2 * M + 3];
        // ---- forward (factored form, as gnn_body)
        for (int side = 0; side < 2; side++) {
            const SideDev &S = side ? a.Z : a.X;
            const int oW1 = side ? Lay::W1z : Lay::W1x, ob1 = side ? Lay::b1z : Lay::b1x;
            const int oW2 = side ? Lay::W2z : Lay::W2x, ob2 = side ? Lay::b2z : Lay::b2x;
            const View2<const float> &logit = side ? a.logit_hz : a.logit_hx;
            const View2<const uint8_t> &synd = side ? a.sz : a.sx;
            float *ft = side ? a.ftz : a.ftx;
            const int e0 = S.vn_ptr[v], e1 = S.vn_ptr[v + 1];
            const float dg = (float)(e1 - e0);
            float *hs1 = a.hs1 + ((int64_t)side * a.B * n + r) * (H + 1);
            float red[M];
            for (int i = 0; i < M; i++) red[i] = 0.0f;
            for (int j = 0; j < H; j++) {
                const float base = FB_FMA(f[2], w[oW1 + 3 * H + j], FB_FMA(f[1], w[oW1 + 2 * H + j], FB_FMA(f[0], w[oW1 + H + j], 0.0f)));
                float hs = 0.0f;
                for (int e = e0; e < e1; e++) {
                    const int c = S.vn_cn[e];
                    const float lg = logit(c, b);
                    const float hc = synd(c, b) ? -lg : lg;
                    const float t = MathExact::tanh(FB_ADD(FB_FMA(hc, w[oW1 + j], base), w[ob1 + j]));
                    hs = (e == e0) ? t : FB_ADD(hs, t);
                    if (j == 0) {
                        float *fr = ft + (b * S.E + e) * 5;
                        fr[0] = hc; fr[1] = f[0]; fr[2] = f[1]; fr[3] = f[2]; fr[4] = 1.0f;
                    }
                }
                for (int i = 0; i < M; i++) red[i] = FB_FMA(hs, w[oW2 + j * M + i], red[i]);
                hs1[j] = (a.reduce == 0 && e1 > e0) ? hs / dg : hs;
            }
            hs1[H] = (e1 > e0) ? ((a.reduce == 0) ? 1.0f : dg) : 0.0f;
            for (int i = 0; i < M; i++) {
                float x = red[i];
                if (a.reduce == 0) x = FB_ADD(FB_DIV(x, dg), w[ob2 + i]);
                else x = FB_FMA(dg, w[ob2 + i], x);
                in[side * M + i] = (e1 > e0) ? x : 0.0f;
            }
        }
        in[2 * M] = f[0]; in[2 * M + 1] = f[1]; in[2 * M + 2] = f[2];
        float *in1 = a.in1 + r * (2 * M + 4);
        for (int k = 0; k < 2 * M + 3; k++) in1[k] = in[k];
        in1[2 * M + 3] = 1.0f;
        // ---- reverse
        const float d0 = a.dout[(b * 3 + 0) * n + v], d1 = a.dout[(b * 3 + 1) * n + v], d2 = a.dout[(b * 3 + 2) * n + v];
        a.dout_r[r * 3 + 0] = d0; a.dout_r[r * 3 + 1] = d1; a.dout_r[r * 3 + 2] = d2;
        float din[2 * M];
        for (int k = 0; k < 2 * M; k++) din[k] = 0.0f;
        float *hid1 = a.hid1 + r * (H + 1), *dpre3 = a.dpre3 + r * H;
        for (int j = 0; j < H; j++) {
            float acc = 0.0f;
            for (int k = 0; k < 2 * M + 3; k++) acc = FB_FMA(in[k], w[Lay::W3 + k * H + j], acc);
            const float hj = MathExact::tanh(FB_ADD(acc, w[Lay::b3 + j]));
            hid1[j] = hj;
            const float dh = d0 * w[Lay::W0 + j * 3 + 0] + d1 * w[Lay::W0 + j * 3 + 1] + d2 * w[Lay::W0 + j * 3 + 2];
            const float dp = dh * (1.0f - hj * hj);
            dpre3[j] = dp;
            for (int k = 0; k < 2 * M; k++) din[k] += w[Lay::W3 + k * H + j] * dp;
        }
        hid1[H] = 1.0f;
        for (int side = 0; side < 2; side++) {
            const SideDev &S = side ? a.Z : a.X;
            const int oW1 = side ? Lay::W1z : Lay::W1x, ob1 = side ? Lay::b1z : Lay::b1x, oW2 = side ? Lay::W2z : Lay::W2x;
            const View2<const float> &logit = side ? a.logit_hz : a.logit_hx;
            const View2<const uint8_t> &synd = side ? a.sz : a.sx;
            float *dp1 = side ? a.dpz : a.dpx;
            const int e0 = S.vn_ptr[v], e1 = S.vn_ptr[v + 1];
            const float scale = (a.reduce == 0 && e1 > e0) ? 1.0f / (float)(e1 - e0) : 1.0f;
            float *dm = a.dm + ((int64_t)side * a.B * n + r) * M;
            for (int i = 0; i < M; i++) dm[i] = (e1 > e0) ? din[side * M + i] : 0.0f;
            for (int j = 0; j < H; j++) {
                float dhs = 0.0f;
                for (int i = 0; i < M; i++) dhs += w[oW2 + j * M + i] * din[side * M + i];
                dhs *= scale;
                const float base = FB_FMA(f[2], w[oW1 + 3 * H + j], FB_FMA(f[1], w[oW1 + 2 * H + j], FB_FMA(f[0], w[oW1 + H + j], 0.0f)));
                for (int e = e0; e < e1; e++) {
                    const int c = S.vn_cn[e];
                    const float lg = logit(c, b);
                    const float hc = synd(c, b) ? -lg : lg;
                    const float t = MathExact::tanh(FB_ADD(FB_FMA(hc, w[oW1 + j], base), w[ob1 + j]));
                    dp1[(b * S.E + e) * H + j] = dhs * (1.0f - t * t);
                }
            }
        }
    }
}

// C_partial[blockIdx] = A[rows of this block]^T B[rows of this block];  A [R][Ka], B [R][Kb], Ka*Kb <= 256*8
__global__ void __launch_bounds__(256) k_atb_partial(const float *__restrict__ A, const float *__restrict__ Bm, int64_t R, int Ka,
                                                     int Kb, float *__restrict__ partial) {
    extern __shared__ float tsm[];
    constexpr int TR = 32, MAXO = 8;
    float *As = tsm, *Bs = tsm + TR * Ka;
    const int nout = Ka * Kb, tid = threadIdx.x;
    float acc[MAXO];
#pragma unroll
    for (int o = 0; o < MAXO; o++) acc[o] = 0.0f;
    const int64_t per = (R + gridDim.x - 1) / gridDim.x, r0 = (int64_t)blockIdx.x * per, r1 = min(R, r0 + per);
    for (int64_t rb = r0; rb < r1; rb += TR) {
        const int nr = (int)min((int64_t)TR, r1 - rb);
        __syncthreads();
        for (int i = tid; i < nr * Ka; i += 256) As[i] = A[rb * Ka + i];
        for (int i = tid; i < nr * Kb; i += 256) Bs[i] = Bm[rb * Kb + i];
        __syncthreads();
#pragma unroll
        for (int o = 0; o < MAXO; o++) {
            const int idx = tid + o * 256;
            if (idx < nout) {
                const int i = idx / Kb, j = idx - i * Kb;
                float s = acc[o];
                for (int r = 0; r < nr; r++) s += As[r * Ka + i] * Bs[r * Kb + j];
                acc[o] = s;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < MAXO; o++) {
        const int idx = tid + o * 256;
        if (idx < nout) partial[(int64_t)blockIdx.x * nout + idx] = acc[o];
    }
}

__global__ void k_atb_reduce(const float *__restrict__ partial, int nblocks, int nout, float *__restrict__ C) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nout) return;
    float s = 0.0f;
    for (int p = 0; p < nblocks; p++) s += partial[(int64_t)p * nout + idx];
    C[idx] = s;
}

}  // namespace train
}  // namespace fbgnn

#endif  // FBGNN_TRAIN_CUH
