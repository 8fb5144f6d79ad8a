"""GPU tests at BASELINE.json's full sizes.  The oracle cannot produce 10^5 frames in seconds, so
these use size-independent properties -- recomputing flags from the residual errors with the
reference's dense formulation, shard/batch invariance of the frame-id addressed sampling, result
identity of round skipping -- and the reference's published logical error rates (binomial tests)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(F, code, wfile, nG, **kw):
    import os
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    F.load_weights(G, os.path.join(F.WEIGHTS_DIR, wfile))
    d1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    return F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [G] * nG, num_layers=nG + 1, **kw)


W1270 = "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy"
W882 = "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"


def _gf2_any(rows, vecs):
    """any((rows @ vecs.T) mod 2) per vector, via float32 BLAS (exact: sums < 2^24)."""
    prod = vecs.astype(np.float32) @ rows.T.astype(np.float32)
    return np.any(prod.astype(np.int64) & 1, axis=1)


def test_config2_full_batch_flags_match_dense_reference_formulation(c1270):
    """configs[2]: B = 5000 (the reference's batch), nG = 3.  Flags recomputed on the host exactly as
    feedback_gnn.py:346-359 builds s_hat / ls_hat (dense hx_perp / hz_perp products)."""
    import fbgnn as F
    B, p = 5000, 0.12
    model = _model(F, c1270, W1270, 3, seed=21)
    res = model.run(B, p, want_counters=True)
    flags = res["flags"].numpy()
    xd, zd = res["x_diff"].numpy(), res["z_diff"].numpy()
    flagged = _gf2_any(c1270.hz, xd) | _gf2_any(c1270.hx, zd)
    block = _gf2_any(c1270.hx_perp, xd) | _gf2_any(c1270.hz_perp, zd)
    assert np.array_equal(flags & 1, flagged.astype(np.uint8))
    assert np.array_equal((flags >> 1) & 1, block.astype(np.uint8))
    assert res["counters"].tolist() == [B, int(flagged.sum()), int(block.sum()), int(((flags >> 2) > 0).sum())]
    # frames that never failed a stage carry rounds == 0; rounds never exceed nG
    assert (flags >> 2).max() <= 3
    # round skipping is result-identical
    model2 = _model(F, c1270, W1270, 3, seed=21, skip_inactive=True)
    res2 = model2.run(B, p, want_counters=True)
    assert np.array_equal(res2["flags"].numpy(), flags) and np.array_equal(res2["x_diff"].numpy(), xd)
    # batching / sharding invariance: two shards of 2500 == one batch of 5000
    a = _model(F, c1270, W1270, 3, seed=21, first_frame=0).run(2500, p)["flags"].numpy()
    b = _model(F, c1270, W1270, 3, seed=21, first_frame=2500).run(2500, p)["flags"].numpy()
    assert np.array_equal(np.concatenate([a, b]), flags)


def test_config1_binary_bp_both_sides_batch_1e5(c1270):
    """configs[1]: [[1270,28]] binary syndrome BP on the X and Z parts separately, batch 10^5."""
    import fbgnn as F
    B, p = 100000, 0.06
    pb = 2 * p / 3
    for pcm, logical, seed in ((c1270.hx, c1270.hz_perp, 31), (c1270.hz, c1270.hx_perp, 32)):
        dec = F.LDPCBPDecoder(pcm, is_syndrome=True, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi")
        model = F.BP_BSC_Model(pcm=pcm, decoder=dec, logical_pcm=logical, p0=0.2, seed=seed)
        res = model.run(B, pb, want_counters=True)
        flags = res["flags"].numpy()
        c = res["counters"]
        assert c[0] == B and c[1] == int((flags & 1).sum()) and c[2] == int(((flags >> 1) & 1).sum())
        assert np.all(((flags >> 1) & 1) >= (flags & 1))            # flagged implies block error
        # determinism + shard invariance
        model_b = F.BP_BSC_Model(pcm=pcm, decoder=dec, logical_pcm=logical, p0=0.2, seed=seed, first_frame=60000)
        assert np.array_equal(model_b.run(40000, pb)["flags"].numpy(), flags[60000:])
        # the layer-by-layer path agrees with the fused one on a slice
        ctx = F.default_context()
        from fbgnn import _ffi
        g = dec.graph()
        noise = ctx.empty((2000, 1270), np.uint8)
        _ffi.call("fbgnn_bsc_sample", ctx.handle, 1270, 2000, float(np.float32(pb)), seed, 0, noise.t2())
        synd = ctx.empty((2000, g.m), np.uint8).T
        _ffi.call("fbgnn_syndrome", g.handle, 2000, noise.t2(), synd.t2())
        llr = np.full((2000, 1270), model.llr_const(pb), np.float32)
        hard = dec((ctx.asarray(llr), synd)).numpy()
        diff = hard ^ noise.numpy()
        fl = _gf2_any(pcm, diff)
        blk = fl | _gf2_any(logical, diff)
        assert np.array_equal(flags[:2000] & 1, fl.astype(np.uint8))
        assert np.array_equal((flags[:2000] >> 1) & 1, blk.astype(np.uint8))
        assert 0.001 < c[2] / B < 0.2


def _compatible(k, n, k_pub, n_pub, z=3.7):
    p_pool = (k + k_pub) / (n + n_pub)
    sigma = np.sqrt(p_pool * (1 - p_pool) * (1 / n + 1 / n_pub))
    return abs(k / n - k_pub / n_pub) < z * sigma + 1e-12


@pytest.mark.parametrize("p,k_pub,n_pub,frames", [(0.14, 1986, 5000, 20000), (0.13, 705, 5000, 20000),
                                                  (0.12, 139, 5000, 40000), (0.11, 106, 25000, 100000),
                                                  (0.10, 100, 275000, 600000)])
def test_published_ler_c1270_three_rounds(c1270, p, k_pub, n_pub, frames):
    """examples/n1270.ipynb cell 2: (64,G,16,G,16,G,16), f=1.0, p0=0.05 -- block errors / frames."""
    import fbgnn as F
    model = _model(F, c1270, W1270, 3, seed=1000 + int(p * 100), skip_inactive=True)
    k = 0
    for _ in range(frames // 20000):
        k += int(model.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"][2])
    assert _compatible(k, frames, k_pub, n_pub), (p, k, frames, k_pub, n_pub)


@pytest.mark.parametrize("p,bler_pub,frames", [(0.12, 6.74e-2, 20000), (0.11, 1.52e-2, 40000), (0.10, 2.40e-3, 200000)])
def test_published_ler_c882_five_rounds(codes, p, bler_pub, frames):
    """examples/n882.ipynb cell 3 (= n882.py, nG = 5): BLER 6.74e-2, 1.52e-2, 2.40e-3 at p = 0.12, 0.11, 0.10;
    each published point holds >= 100 block errors."""
    import fbgnn as F
    model = _model(F, codes["c882"], W882, 5, seed=2000 + int(p * 100), skip_inactive=True)
    k = 0
    for _ in range(frames // 20000):
        k += int(model.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"][2])
    n_pub = max(int(round(100 / bler_pub)), 5000)
    assert _compatible(k, frames, int(round(bler_pub * n_pub)), n_pub), (p, k, frames)


def test_sim_ber_drop_in_flow(codes):
    """The n882.py flow end to end: PlotBER.simulate(model, qldpc=True) stops on the target and the
    BLER curve sits at index 1."""
    import fbgnn as F
    model = _model(F, codes["c882"], W882, 5, seed=5)
    plot = F.PlotBER()
    plot.simulate(model, ebno_dbs=[0.12], batch_size=5000, num_target_block_errors=100, legend="feedback GNN 1.00 5 rounds",
                  soft_estimates=True, max_mc_iter=100, early_stop=True, add_bler=True, show_fig=False, qldpc=True,
                  forward_keyboard_interrupt=False, verbose=False)
    bler = float(plot._bers[1][0])
    assert 0.04 < bler < 0.10 and plot._is_bler == [False, True]
    assert F.sim_ber.last["nb_blocks"][0] == 5000 and F.sim_ber.last["status"][0] == 4
