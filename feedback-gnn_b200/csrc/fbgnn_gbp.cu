// fbgnn_gbp.cu -- GNN_BP4 (gnn.py:71-751, BASELINE configs[4]): weight packing, kernel sequencing, C ABI.
#include "fbgnn_internal.h"
#include "fbgnn_gbp_tc.cuh"

// ------------------------------------------------------------------ GNN_BP4 -------------
struct fbgnn_gbp {
    fbgnn_ctx *ctx;
    int d, H, M, act, reduce, use_bias;
    int gemm = FBGNN_GEMM_FMA;
    int64_t chunk_limit = 0;       // > 0: at most this many frames per pass (fbgnn_gbp_set_chunk)
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

extern "C" int fbgnn_gbp_set_chunk(fbgnn_gbp *g, int64_t max_frames) {
    REQUIRE(g, "NULL handle");
    REQUIRE(max_frames >= 0, "max_frames must be non-negative (0 = automatic)");
    g->chunk_limit = max_frames;
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
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)6 << 30) / per_frame));
    if (g->chunk_limit > 0) chunk = std::min(chunk, g->chunk_limit);
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
        rc = ctx->math_mode == FBGNN_MATH_SFU ? gbp_run<MathSfu>(code, g, num_iter, nb, a, xl, zl, xh, zh)
                                               : gbp_run<MathExact>(code, g, num_iter, nb, a, xl, zl, xh, zh);
    }
    CK(cudaFreeAsync(h_vn, st)); CK(cudaFreeAsync(hcx, st)); CK(cudaFreeAsync(hcz, st)); CK(cudaFreeAsync(lg, st));
    if (fact) { CK(cudaFreeAsync(pfc, st)); CK(cudaFreeAsync(pfv, st)); }
    return rc;
}

