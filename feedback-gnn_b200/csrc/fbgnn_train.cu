// fbgnn_train.cu -- second training stage: loss and weight gradients (C ABI fbgnn_second_stage_grad).
#include "fbgnn_internal.h"
#include "fbgnn_train.cuh"

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
    if (!(gnn->layers == 2 && gnn->H == 40 && gnn->M == 20 && gnn->act == FBGNN_ACT_TANH && gnn->use_bias && gnn->reduce <= 1))
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
    // the reverse sweep below differentiates the exact arithmetic: run the forward GNN in it as well
    const int saved_math = ctx->math_mode;
    ctx->math_mode = FBGNN_MATH_EXACT;
    int rc = launch_gnn(ctx, gnn, ga);
    ctx->math_mode = saved_math;

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

