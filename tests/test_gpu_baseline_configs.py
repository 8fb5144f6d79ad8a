"""GPU parity on the BASELINE.json configurations themselves ("parity holes"): the headline code
[[1270,28]] at the layer level, the nG = 3 pipeline frame by frame, configs[0] literally, the chunked GNN_BP4
path that produces the configs[4] number, and the fixed-point exit.  Everything goes through the C ABI and is
compared BIT FOR BIT with the CPU oracle on identical inputs (the two share the arithmetic specification
fb_math.h; the numpy oracle guards that header in tests/test_oracle.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bitexact(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = (_bits(a) == _bits(b)) if a.dtype.kind == "f" else (a == b)
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} entries differ"


def _noise_and_syndromes(oracle, code, B, p, seed, first_frame=0):
    nx, nz = oracle.pauli(seed, first_frame, B, code.N, p)
    sx = (code.hx @ nz.T.astype(np.int64)) & 1
    sz = (code.hz @ nx.T.astype(np.int64)) & 1
    return nx, nz, sx.astype(np.uint8), sz.astype(np.uint8)


def _weights_list(W, use_bias=True):
    from oracle.c_oracle import GBP_KEYS
    out = [W["Winv"], W["binv"]]
    for k in GBP_KEYS:
        out += list(W[k])
    return out if use_bias else out[0::2]


def _syndromes(oracle, code, B, p, seed):
    """batch-first syndromes as GNN_BP4 takes them (gnn.py:385-386)"""
    _, _, sx, sz = _noise_and_syndromes(oracle, code, B, p, seed)
    return np.ascontiguousarray(sx.T), np.ascontiguousarray(sz.T)


OUT_KEYS = ("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit", "msg_x", "msg_z")


def _compare_layer(F, oracle, code, g, llr, prior, sx, sz, it, factor, tag):
    """QLDPCBPDecoder.call (decoding_q.py:661-797) through fbgnn_bp4_decode vs oracle.bp4: marginals, decisions,
    both soft syndromes and the final check-to-variable messages."""
    dec = F.QLDPCBPDecoder(code, num_iter=it, normalization_factor=factor, cn_type="boxplus-phi", stage_one=True)
    ctx = dec._device().ctx
    d_llr = None if llr is None else ctx.asarray(llr)
    out = dec.decode_device(d_llr, ctx.asarray(sx), ctx.asarray(sz), want_logits=True, want_msgs=True,
                            prior=None if llr is not None else float(prior))
    ref = oracle.bp4(g, float(prior) if llr is None else llr, sx, sz, it, factor, "boxplus-phi", want_msgs=True)
    for k, o in zip(OUT_KEYS, out):
        assert_bitexact(o.numpy(), ref[k], f"{tag} it={it} {k}")
    return ref


@pytest.mark.parametrize("it", [1, 16, 64])
@pytest.mark.parametrize("prior_kind", ["const", "per_variable"])
def test_bp4_layer_bitexact_c1270(c1270, oracle, it, prior_kind):
    """(a) the headline code at the layer level, constant prior (stage 0 of the pipeline) and per-variable priors
    (the later stages), 1 / 16 / 64 iterations -- the 64-iteration case runs the fixed-point-exit kernel."""
    import fbgnn as F
    B = 96
    _, _, sx, sz = _noise_and_syndromes(oracle, c1270, B, 0.10, seed=41)
    g = oracle.CodeGraph(c1270)
    prior = oracle.prior_llr(0.05)
    llr = None
    if prior_kind == "per_variable":
        rng = np.random.default_rng(17)
        llr = (prior + rng.normal(0, 0.4, (B, 3, c1270.N))).astype(np.float32)
    _compare_layer(F, oracle, c1270, g, llr, prior, sx, sz, it, 1.0, f"c1270 {prior_kind}")


@pytest.mark.parametrize("p", [0.12, 0.10])
@pytest.mark.parametrize("skip", [False, True])
def test_pipeline_bitexact_c1270_three_rounds(c1270, oracle, weights, p, skip):
    """(b) BASELINE configs[2] frame by frame: (64, G, 16, G, 16, G, 16) with the shipped weights on 256 frames;
    flags (incl. the number of rounds a frame stayed active), residual errors and counters
    (feedback_gnn.py:293-361)."""
    import fbgnn as F
    B = 256
    G = F.Feedback_GNN(code=c1270, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    G.set_weights(weights["c1270"])
    d1 = F.QLDPCBPDecoder(c1270, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(c1270, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(c1270, [d1] + [d2] * 3, [G] * 3, num_layers=4, p0=0.05, seed=2,
                                               first_frame=5000, skip_inactive=skip)
    res = model.run(B, p, want_counters=True)
    ref = oracle.pipeline(oracle.CodeGraph(c1270), [64, 16, 16, 16], [oracle.Gnn(weights["c1270"])] * 3, p, p0=0.05,
                          seed=2, first_frame=5000, B=B, skip_inactive=False, want_diff=True)
    assert_bitexact(res["flags"].numpy(), ref["flags"], f"c1270 nG=3 p={p} flags")
    assert_bitexact(res["x_diff"].numpy(), ref["x_diff"], "x_diff")
    assert_bitexact(res["z_diff"].numpy(), ref["z_diff"], "z_diff")
    assert res["counters"].tolist() == ref["counters"].tolist()
    assert ref["counters"][3] > 0                       # some frames do fail stage 0, so the GNN rounds matter


@pytest.mark.parametrize("p", [0.01, 0.05, 0.10])
def test_config0_literally(codes, oracle, p):
    """(c) BASELINE configs[0]: [[882,24]], quaternary BP, batch 1000, 32 iterations, normalisation 0.625
    (the decoder defaults of decoding_q.py:21-22 as QLDPC.ipynb cell 11 uses them), prior p0 = p."""
    import fbgnn as F
    code = codes["c882"]
    B = 1000
    _, _, sx, sz = _noise_and_syndromes(oracle, code, B, p, seed=0)
    prior = oracle.prior_llr(p)
    llr = np.full((B, 3, code.N), prior, np.float32)
    dec = F.QLDPCBPDecoder(code, num_iter=32, normalization_factor=0.625, cn_type="boxplus-phi", stage_one=True)
    out = dec((llr, sx, sz))
    ref = oracle.bp4(oracle.CodeGraph(code), llr, sx, sz, 32, 0.625, "boxplus-phi")
    for k, o in zip(OUT_KEYS[:7], out):
        assert_bitexact(np.asarray(o, dtype=ref[k].dtype), ref[k], f"configs[0] p={p} {k}")
    # the constant-prior kernel gives the same bits as the per-variable one fed with a constant
    ctx = dec._device().ctx
    out_c = dec.decode_device(None, ctx.asarray(sx), ctx.asarray(sz), want_logits=True, prior=float(prior))
    for k, o in zip(OUT_KEYS[:7], out_c):
        assert_bitexact(o.numpy(), ref[k], f"configs[0] const-prior p={p} {k}")


@pytest.mark.parametrize("gemm", ["fma", "tf32x3"])
def test_gnn_bp4_chunked_batches(oracle, codes, gemm):
    """(d) BASELINE configs[4] runs its 65 536 frames in several passes over the batch (fbgnn_gbp_decode's chunk
    loop).  Force 4 chunks with a ragged tail (B = 53, 16 frames per pass) and compare with the one-pass oracle:
    FMA bit-exact, tensor-core path within its stated tolerance (gnn.py:385-420)."""
    import fbgnn as F
    code = codes["c882"]
    W = oracle.gnn_bp4_random_weights(seed=4, use_bias=True)
    B, it = 53, 3
    sx, sz = _syndromes(oracle, code, B, 0.06, 19)
    ref = oracle.gnn_bp4(oracle.CodeGraph(code), W, sx, sz, it, reduce_op="mean")
    G = F.GNN_BP4(code, num_embed_dims=20, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, num_iter=it,
                  reduce_op="mean", activation="tanh", use_bias=True, gemm=gemm, chunk_frames=16)
    G.set_weights(_weights_list(W, True))
    llr_hat, x_hat, z_hat = G((sx, sz))
    for i in range(it):
        for got, want in ((llr_hat[i][0], ref["x_logit"][i]), (llr_hat[i][1], ref["z_logit"][i])):
            if gemm == "fma":
                assert_bitexact(got, want, f"chunked GNN_BP4 it={i}")
            else:
                assert np.all(np.abs(got - want) <= 2e-5 + 1e-5 * np.abs(want)), float(np.abs(got - want).max())
    agree = np.mean((x_hat.astype(np.uint8) == ref["x_hat"]) & (z_hat.astype(np.uint8) == ref["z_hat"]))
    assert agree == 1.0 if gemm == "fma" else agree >= 0.9999
    # and the chunked result equals the automatic (single-chunk) one bit for bit on the FMA path
    if gemm == "fma":
        G1 = F.GNN_BP4(code, 20, 20, 40, 2, it, reduce_op="mean", activation="tanh", use_bias=True)
        G1.set_weights(_weights_list(W, True))
        one, xh1, zh1 = G1((sx, sz))
        for i in range(it):
            assert_bitexact(one[i][0], llr_hat[i][0], "chunked == unchunked x_logit")
            assert_bitexact(one[i][1], llr_hat[i][1], "chunked == unchunked z_logit")


@pytest.mark.parametrize("name,p", [("c1270", 0.02), ("c882", 0.02), ("c1270", 0.06)])
def test_fixed_point_exit_is_exact(codes, c1270, oracle, name, p):
    """(e) 64-iteration runs at low p: nearly every frame reaches a bit-exact fixed point long before iteration 64
    and the kernel leaves its loop early; the oracle runs all 64 iterations.  Messages, marginals, soft syndromes
    and decisions must still agree bit for bit, and most frames must indeed have converged (so the exit path is
    what is being tested)."""
    import fbgnn as F
    code = c1270 if name == "c1270" else codes[name]
    B = 192
    _, _, sx, sz = _noise_and_syndromes(oracle, code, B, p, seed=77)
    g = oracle.CodeGraph(code)
    prior = oracle.prior_llr(0.05)
    ref = _compare_layer(F, oracle, code, g, None, prior, sx, sz, 64, 1.0, f"{name} fpx p={p}")
    # converged = the decision reproduces the syndrome
    ok = np.all(((code.hx @ ref["z_hat"].T.astype(np.int64)) & 1) == sx, axis=0) & \
         np.all(((code.hz @ ref["x_hat"].T.astype(np.int64)) & 1) == sz, axis=0)
    if p <= 0.02:
        assert ok.mean() > 0.9
    # 63 vs 64 iterations differ for frames that have NOT reached a fixed point -- the exit is not a blanket skip
    ref63 = oracle.bp4(g, float(prior), sx, sz, 63, 1.0, "boxplus-phi", want_msgs=True)
    if p > 0.02:
        assert not np.array_equal(_bits(ref63["msg_x"]), _bits(ref["msg_x"]))
    rng = np.random.default_rng(5)
    llr = (prior + rng.normal(0, 0.2, (B, 3, code.N))).astype(np.float32)
    _compare_layer(F, oracle, code, g, llr, prior, sx, sz, 64, 1.0, f"{name} fpx per-variable p={p}")


def test_gnn_layer_bitexact_c1270(c1270, oracle, weights):
    """Feedback_GNN.call (feedback_gnn.py:161-188) on the headline code with both shipped weight sets, fed with the
    marginals and soft syndromes of a 64-iteration first stage."""
    import fbgnn as F
    B = 64
    _, _, sx, sz = _noise_and_syndromes(oracle, c1270, B, 0.12, seed=6)
    g = oracle.CodeGraph(c1270)
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 64)
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
    for wkey in ("c1270", "c1270_coarse"):
        G = F.Feedback_GNN(code=c1270, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                           activation="tanh", use_bias=True)
        G.set_weights(weights[wkey])
        out = G((h_vn, r["z_logit"], r["x_logit"], sx, sz))
        ref = oracle.gnn(g, oracle.Gnn(weights[wkey], "tanh", "mean"), h_vn, r["z_logit"], r["x_logit"], sx, sz)
        assert_bitexact(out, ref, f"gnn c1270 {wkey}")


def test_bp2_layer_bitexact_c1270(c1270, oracle):
    """BASELINE configs[1] at the layer level: binary syndrome BP on hx (Z part) and hz (X part), 64 iterations."""
    import fbgnn as F
    B, pb = 128, 2 * 0.08 / 3
    for pcm, seed in ((c1270.hx, 31), (c1270.hz, 32)):
        noise = oracle.bsc(seed, 0, B, c1270.N, pb)
        synd = ((pcm @ noise.T.astype(np.int64)) & 1).astype(np.uint8)
        llr = np.full((B, c1270.N), -np.log((1 - 0.2) / 0.2), np.float32)
        soft_ref, hard_ref = oracle.bp2(pcm, llr, synd, 64, 1.0, "boxplus-phi")
        dec_s = F.LDPCBPDecoder(pcm, is_syndrome=True, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi",
                                hard_out=False)
        dec_h = F.LDPCBPDecoder(pcm, is_syndrome=True, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi")
        assert_bitexact(dec_s((llr, synd)), soft_ref, "bp2 c1270 soft")
        assert_bitexact(dec_h((llr, synd)).astype(np.uint8), hard_ref, "bp2 c1270 hard")
