"""GPU parity tests proper: the CUDA path (through the C ABI via the fbgnn layers) against the
CPU oracle on identical inputs.  Integer outputs (syndromes, decisions, flags, counters) and --
because the kernels and the oracle share the arithmetic specification fb_math.h -- all float32
messages, marginals, soft syndromes and GNN outputs must be BIT-EXACT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bitexact(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        same = _bits(a) == _bits(b)
    else:
        same = a == b
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} entries differ"


def _noise_and_syndromes(oracle, code, B, p, seed):
    n = code.hx.shape[1]
    nx, nz = oracle.pauli(seed, 0, B, n, p)
    sx = (code.hx @ nz.T.astype(np.int64)) & 1
    sz = (code.hz @ nx.T.astype(np.int64)) & 1
    return nx, nz, sx.astype(np.uint8), sz.astype(np.uint8)


@pytest.mark.parametrize("name,B,p", [("steane", 64, 0.05), ("rsurf3", 64, 0.08), ("toric4", 48, 0.05),
                                      ("gb48", 64, 0.05), ("c882", 96, 0.09)])
@pytest.mark.parametrize("cn_type,factor", [("boxplus-phi", 1.0), ("boxplus-phi", 0.625), ("minsum", 0.8),
                                            ("boxplus", 1.0)])
def test_bp4_layer_bitexact(codes, oracle, name, B, p, cn_type, factor):
    import fbgnn as F
    code = codes[name]
    n = code.N
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, p, seed=11)
    rng = np.random.default_rng(3)
    prior = oracle.prior_llr(0.05)
    llr = (prior + rng.normal(0, 0.3, (B, 3, n))).astype(np.float32)
    g = oracle.CodeGraph(code)
    for it in (0, 1, 2, 7, 32):
        dec = F.QLDPCBPDecoder(code, num_iter=it, normalization_factor=factor, cn_type=cn_type, stage_one=True)
        out = dec((llr, sx, sz))
        ref = oracle.bp4(g, llr, sx, sz, it, factor, cn_type, want_msgs=True)
        for k, o in zip(("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit"), out):
            assert_bitexact(np.asarray(o, dtype=ref[k].dtype), ref[k], f"{name} {cn_type} it={it} {k}")
        dev = dec._device()
        d = dec.decode_device(dev.ctx.asarray(llr), dev.ctx.asarray(sx), dev.ctx.asarray(sz), want_msgs=True)
        assert_bitexact(d[7].numpy(), ref["msg_x"], f"{name} {cn_type} it={it} msg_x")
        assert_bitexact(d[8].numpy(), ref["msg_z"], f"{name} {cn_type} it={it} msg_z")


@pytest.mark.parametrize("name", ["c882", "toric4"])
def test_bp4_stage_two_per_iteration_soft_syndromes(codes, oracle, name):
    """stage_two=True returns (llr_hat [2*num_iter+2, m, B], x_hat, z_hat) (decoding_q.py:794-795)."""
    import fbgnn as F
    code = codes[name]
    B, it = 24, 6
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.08, seed=13)
    rng = np.random.default_rng(5)
    llr = (oracle.prior_llr(0.05) + rng.normal(0, 0.2, (B, 3, code.N))).astype(np.float32)
    dec = F.QLDPCBPDecoder(code, num_iter=it, normalization_factor=1.0, cn_type="boxplus-phi", stage_two=True)
    llr_hat, x_hat, z_hat = dec((llr, sx, sz))
    ref = oracle.bp4(oracle.CodeGraph(code), llr, sx, sz, it, 1.0, "boxplus-phi", want_iter_logits=True)
    assert llr_hat.shape == (2 * it + 2, code.hx.shape[0], B)
    assert_bitexact(llr_hat, ref["llr_hat"], f"{name} llr_hat")
    assert_bitexact(x_hat.astype(np.uint8), ref["x_hat"], "x_hat")
    assert_bitexact(z_hat.astype(np.uint8), ref["z_hat"], "z_hat")
    assert_bitexact(llr_hat[2 * it], ref["x_logit"], "last slot == x_logit")


def test_bp4_output_dtypes_and_plain_mode(codes, oracle):
    import fbgnn as F
    code = codes["c882"]
    B = 8
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.05, seed=2)
    llr = np.full((B, 3, code.N), oracle.prior_llr(0.05), np.float32)
    out = F.QLDPCBPDecoder(code, num_iter=8, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)((llr, sx, sz))
    assert [o.dtype for o in out] == [np.float32] * 3 + [np.int64, np.float64, np.float32, np.float32]
    assert out[5].shape == (code.hz.shape[0], B) and out[6].shape == (code.hx.shape[0], B)
    xh, zh = F.QLDPCBPDecoder(code, num_iter=8, normalization_factor=1.0, cn_type="boxplus-phi")((llr, sx, sz))
    assert np.array_equal(xh, out[3]) and np.array_equal(zh, out[4])
    with pytest.raises(TypeError):
        F.QLDPCBPDecoder(code, num_iter=1)((llr.astype(np.float64), sx, sz))
    with pytest.raises(ValueError):
        F.QLDPCBPDecoder(code, num_iter=1)((llr[:, :, :-1], sx, sz))


@pytest.mark.parametrize("name", ["steane", "rsurf3", "gb48", "c882"])
@pytest.mark.parametrize("cn_type", ["boxplus-phi", "minsum", "boxplus"])
def test_bp2_layer_bitexact(codes, oracle, name, cn_type):
    import fbgnn as F
    code = codes[name]
    B, n = 80, code.N
    noise = oracle.bsc(5, 0, B, n, 0.04)
    synd = ((code.hx @ noise.T.astype(np.int64)) & 1).astype(np.uint8)
    rng = np.random.default_rng(0)
    llr = (-np.log((1 - 0.1) / 0.1) + rng.normal(0, 0.2, (B, n))).astype(np.float32)
    llr[0, :3] = [30.0, -30.0, 0.0]                      # exercises the +-20 clip
    for it in (0, 1, 3, 20):
        soft_ref, hard_ref = oracle.bp2(code.hx, llr, synd, it, 0.9, cn_type)
        hard = F.LDPCBPDecoder(code.hx, is_syndrome=True, num_iter=it, normalization_factor=0.9, cn_type=cn_type)((llr, synd))
        soft = F.LDPCBPDecoder(code.hx, is_syndrome=True, num_iter=it, normalization_factor=0.9, cn_type=cn_type,
                               hard_out=False)((llr, synd))
        assert_bitexact(soft, soft_ref, f"{name} {cn_type} it={it} soft")
        assert_bitexact(hard.astype(np.uint8), hard_ref, f"{name} {cn_type} it={it} hard")
    # non-syndrome mode (plain codeword decoding) goes through the same kernel
    soft_ref, _ = oracle.bp2(code.hx, llr, None, 5, 1.0, cn_type)
    soft = F.LDPCBPDecoder(code.hx, num_iter=5, cn_type=cn_type, hard_out=False)(llr)
    assert_bitexact(soft, soft_ref, f"{name} {cn_type} no-syndrome")


@pytest.mark.parametrize("name,wkey", [("c882", "c882"), ("c882", "c882_coarse")])
@pytest.mark.parametrize("reduce_op", ["mean", "sum", "max", "min"])
def test_gnn_layer_bitexact(codes, oracle, weights, name, wkey, reduce_op):
    import fbgnn as F
    code = codes[name]
    B = 40
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.1, seed=4)
    g = oracle.CodeGraph(code)
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 16)
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op=reduce_op,
                       activation="tanh", use_bias=True)
    G.set_weights(weights[wkey])
    assert G.count_params() == 3923                      # examples/Feedback_GNN.ipynb cell 6
    out = G((h_vn, r["z_logit"], r["x_logit"], sx, sz))
    ref = oracle.gnn(g, oracle.Gnn(weights[wkey], "tanh", reduce_op), h_vn, r["z_logit"], r["x_logit"], sx, sz)
    assert_bitexact(out, ref, f"gnn {wkey} {reduce_op}")


def test_gnn_irregular_code_and_relu(codes, oracle):
    import fbgnn as F
    code = codes["rsurf3"]
    B = 33
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.1, seed=9)
    g = oracle.CodeGraph(code)
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 5)
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
    for act, bias in (("relu", True), ("tanh", False)):
        G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, activation=act,
                           use_bias=bias)
        w = G.get_weights()
        rng = np.random.default_rng(1)
        w = [(a + rng.normal(0, 0.3, a.shape)).astype(np.float32) for a in w]
        G.set_weights(w)
        out = G((h_vn, r["z_logit"], r["x_logit"], sx, sz))
        ref = oracle.gnn(g, oracle.Gnn(w, act, "mean", use_bias=bias), h_vn, r["z_logit"], r["x_logit"], sx, sz)
        assert_bitexact(out, ref, f"gnn rsurf3 {act}")


@pytest.mark.parametrize("L,reduce_op,bias", [(1, "mean", True), (3, "mean", True), (3, "max", False), (4, "sum", True)])
def test_gnn_any_mlp_depth_bitexact(codes, oracle, L, reduce_op, bias):
    """Feedback_GNN(num_mlp_layers != 2) (feedback_gnn.py:110-127): layer and a 2-stage pipeline vs the oracle."""
    import fbgnn as F
    code = codes["gb48"]
    B = 50
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.08, seed=9)
    g = oracle.CodeGraph(code)
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 6)
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=L, reduce_op=reduce_op,
                       activation="tanh", use_bias=bias)
    rng = np.random.default_rng(L)
    w = [(a + rng.normal(0, 0.25, a.shape)).astype(np.float32) for a in G.get_weights()]
    G.set_weights(w)
    out = G((h_vn, r["z_logit"], r["x_logit"], sx, sz))
    ref = oracle.gnn_deep(g, oracle.GnnDeep(w, 40, 20, L, "tanh", reduce_op, use_bias=bias), h_vn, r["z_logit"],
                          r["x_logit"], sx, sz)
    assert_bitexact(out, ref, f"deep gnn L={L} {reduce_op}")


def test_pauli_and_syndrome_bitexact(codes, oracle):
    import fbgnn as F
    from fbgnn import _ffi
    code = codes["c882"]
    B, n, p = 257, code.N, 0.1
    ch = F.Pauli(seed=123, first_frame=1 << 33)
    nx, nz = ch([np.zeros((B, n), np.float32), None, 2 * p / 3, p / 3, 2 * p / 3])
    rx, rz = oracle.pauli(123, 1 << 33, B, n, p)
    assert nx.dtype == bool and np.array_equal(nx, rx.astype(bool)) and np.array_equal(nz, rz.astype(bool))
    assert abs(nx.mean() - 2 * p / 3) < 0.004 and abs((nx & nz).mean() - p / 3) < 0.003
    g = _ffi.Graph(code.hx)
    ctx = g.ctx
    synd = ctx.empty((B, g.m), np.uint8).T
    _ffi.call("fbgnn_syndrome", g.handle, B, ctx.asarray(rz).t2(), synd.t2())
    assert np.array_equal(synd.numpy(), (code.hx @ rz.T.astype(np.int64)) & 1)


def test_fixed_weight_pauli_and_model(codes, oracle, weights):
    """Pauli(wt=True) (pauli.py:80-96) and Sandwich(..., wt=True): exactly wt errors per frame."""
    import fbgnn as F
    code = codes["c882"]
    B, n, wt = 300, code.N, 37
    ch = F.Pauli(wt=True, seed=11, first_frame=70)
    nx, nz = ch([np.zeros((B, n), np.float32), None, wt])
    rx, rz = oracle.pauli_wt(11, 70, B, n, wt)
    assert np.array_equal(nx, rx.astype(bool)) and np.array_equal(nz, rz.astype(bool))
    assert np.all((nx | nz).sum(1) == wt)                           # exactly wt erroneous qubits
    frac = [(nx & ~nz).sum() / (B * wt), (nx & nz).sum() / (B * wt), (~nx & nz).sum() / (B * wt)]
    assert all(abs(f - 1 / 3) < 0.03 for f in frac)                 # X, Y, Z equally likely
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    G.set_weights(weights["c882"])
    d1 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d1], [G], num_layers=2, wt=True, p0=0.05, seed=11)
    res = model.run(128, 60, want_counters=True)
    ref = oracle.pipeline(oracle.CodeGraph(code), [16, 16], [oracle.Gnn(weights["c882"])], 0.05, p0=0.05, seed=11,
                          B=128, wt=60)
    assert np.array_equal(res["flags"].numpy(), ref["flags"]) and np.array_equal(res["counters"], ref["counters"])


@pytest.mark.parametrize("name,nG,p,B", [("c882", 1, 0.12, 192), ("c882", 3, 0.12, 160), ("rsurf3", 0, 0.08, 300),
                                         ("gb48", 2, 0.06, 128)])
@pytest.mark.parametrize("skip", [False, True])
def test_pipeline_bitexact(codes, oracle, weights, name, nG, p, B, skip):
    """Sandwich model, sampled in-kernel: per-frame flags, residual errors and counters."""
    import fbgnn as F
    code = codes[name]
    w = weights["c882"]
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    G.set_weights(w)
    d1 = F.QLDPCBPDecoder(code, num_iter=24, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=8, normalization_factor=0.9, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [G] * nG, num_layers=nG + 1, seed=7,
                                               first_frame=1000, skip_inactive=skip)
    res = model.run(B, p, want_counters=True)
    g = oracle.CodeGraph(code)
    ref = oracle.pipeline(g, [24] + [8] * nG, [oracle.Gnn(w)] * nG, p, p0=0.05, factors=[1.0] + [0.9] * nG,
                          seed=7, first_frame=1000, B=B, skip_inactive=False, want_diff=True)
    assert_bitexact(res["flags"].numpy(), ref["flags"], f"{name} nG={nG} flags")
    assert_bitexact(res["x_diff"].numpy(), ref["x_diff"], "x_diff")
    assert_bitexact(res["z_diff"].numpy(), ref["z_diff"], "z_diff")
    assert np.array_equal(res["counters"], ref["counters"])
    assert model.next_frame == 1000 + B


@pytest.mark.parametrize("name,nG", [("c882", 2), ("rsurf3", 0)])
def test_pipeline_packed_bit_io(codes, oracle, weights, name, nG):
    """fbgnn_pipeline_run_bits: noise in / indicators and residual errors out as packed bit-planes (32 per word) give
    exactly what the byte interface gives, for sampled and for given noise."""
    import fbgnn as F
    code = codes[name]
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    G.set_weights(weights["c882"])
    d1 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    mk = lambda: F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] * (nG + 1), [G] * nG, num_layers=nG + 1, seed=6, first_frame=77)
    B, p = 203, 0.12                                   # a ragged last word of frames
    a = mk().run(B, p, want_counters=True)
    b = mk().run_bits(B, p, want_diff=True, want_counters=True)
    flags = a["flags"].numpy()
    planes = F.unpack_bits(b["frame_bits"].numpy(), B)
    assert np.array_equal(planes[0], flags & 1) and np.array_equal(planes[1], (flags >> 1) & 1)
    assert np.array_equal(planes[2], ((flags >> 2) > 0).astype(np.uint8))
    assert np.array_equal(F.unpack_bits(b["x_diff_bits"].numpy(), code.N), a["x_diff"].numpy())
    assert np.array_equal(F.unpack_bits(b["z_diff_bits"].numpy(), code.N), a["z_diff"].numpy())
    assert a["counters"].tolist() == b["counters"].tolist()
    nx, nz = oracle.pauli(6, 77, B, code.N, p)
    assert np.array_equal(F.unpack_bits(F.pack_bits(nx), code.N), nx)
    c = mk().run_bits(B, p, noise_bits=(F.pack_bits(nx), F.pack_bits(nz)), want_counters=True)
    assert np.array_equal(c["frame_bits"].numpy(), b["frame_bits"].numpy()) and c["counters"].tolist() == a["counters"].tolist()
    with pytest.raises(ValueError):
        mk().run_bits(B, p, noise_bits=(F.pack_bits(nx)[:, :-1], F.pack_bits(nz)[:, :-1]))


def test_pipeline_model_outputs_match_reference_matrices(codes, oracle, weights):
    """model(batch_size, p) -> (s_hat, ls_hat): dense matrices as the reference builds them
    (feedback_gnn.py:349-359) and the sim_ber counters."""
    import fbgnn as F
    code = codes["c882"]
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    F.load_weights(G, F.WEIGHTS_DIR + "/feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy")
    d1 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d1], [G], num_layers=2, seed=5)
    B, p = 64, 0.13
    s_hat, ls_hat = model(B, p)
    g = oracle.CodeGraph(code)
    ref = oracle.pipeline(g, [16, 16], [oracle.Gnn(weights["c882"])], p, seed=5, B=B, want_diff=True)
    xd, zd = ref["x_diff"].astype(np.int64), ref["z_diff"].astype(np.int64)
    s_ref = np.concatenate([(xd @ code.hz.T) & 1, (zd @ code.hx.T) & 1], 1)
    ls_ref = np.concatenate([(xd @ code.hx_perp.T) & 1, (zd @ code.hz_perp.T) & 1], 1)
    assert np.array_equal(np.asarray(s_hat), s_ref) and np.array_equal(np.asarray(ls_hat), ls_ref)
    assert F.count_block_errors(None, s_hat) == int(np.any(s_ref, 1).sum()) == ref["counters"][1]
    assert F.count_block_errors(None, ls_hat) == int(np.any(ls_ref, 1).sum()) == ref["counters"][2]


def test_pipeline_given_noise_edge_cases(codes, oracle, weights):
    """Zero noise, a single-qubit error, a stabiliser and a logical operator as given noise."""
    import fbgnn as F
    code = codes["c882"]
    n = code.N
    nx = np.zeros((5, n), np.uint8)
    nz = np.zeros((5, n), np.uint8)
    nx[1, 17] = 1
    nz[2, :] = code.hz[3]             # a Z-stabiliser: trivial syndrome, no logical error
    nx[3, :] = code.lx[0]             # a logical X: trivial syndrome, block error without flag
    nx[4, 5] = nz[4, 5] = 1           # a Y error
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    G.set_weights(weights["c882"])
    d1 = F.QLDPCBPDecoder(code, num_iter=32, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d1], [G], num_layers=2)
    res = model.run(5, 0.05, noise=(nx, nz), want_counters=True)
    flags = res["flags"].numpy()
    assert list(flags & 3) == [0, 0, 0, 2, 0]
    g = oracle.CodeGraph(code)
    ref = oracle.pipeline(g, [32, 32], [oracle.Gnn(weights["c882"])], 0.05, B=5, noise=(nx, nz))
    assert np.array_equal(flags, ref["flags"])
    assert model.run(0, 0.05, want_counters=True)["counters"].tolist() == [0, 0, 0, 0]   # empty batch


@pytest.mark.parametrize("name,logical", [("c882", True), ("gb48", False)])
def test_bsc_pipeline_bitexact(codes, oracle, name, logical):
    import fbgnn as F
    code = codes[name]
    B, p = 500, 0.04
    dec = F.LDPCBPDecoder(code.hx, is_syndrome=True, num_iter=20, cn_type="boxplus-phi")
    lp = code.hz_perp if logical else None
    if logical:
        model = F.BP_BSC_Model(pcm=code.hx, decoder=dec, logical_pcm=lp, p0=0.2, seed=3)
        res = model.run(B, p, want_counters=True)
        ref = oracle.bsc_pipeline(code.hx, lp, 20, p, p0=0.2, seed=3, B=B)
        assert np.array_equal(res["flags"].numpy(), ref["flags"])
        assert np.array_equal(res["counters"], ref["counters"])
        s_hat, ls_hat = model(B, p)
        assert s_hat.shape == (B, code.hx.shape[0]) and ls_hat.shape == (B, lp.shape[0])
    else:
        model = F.BP_BSC_Model(pcm=code.hx, decoder=dec, p0=0.2, seed=3)
        noise, noise_hat = model(B, p)
        ref_noise = oracle.bsc(3, 0, B, code.N, p)
        assert np.array_equal(noise, ref_noise.astype(np.float32))
        synd = ((code.hx @ ref_noise.T.astype(np.int64)) & 1).astype(np.uint8)
        llr = np.full((B, code.N), -np.log((1 - 0.2) / 0.2), np.float32)
        _, hard = oracle.bp2(code.hx, llr, synd, 20)
        assert np.array_equal(noise_hat, hard.astype(np.float32))


def test_math_probes_bitexact(oracle):
    """The arithmetic specification itself: device functions == host functions on dense samples."""
    import ctypes as C
    from fbgnn import _ffi
    ctx = _ffi.default_context()
    rng = np.random.default_rng(0)
    cases = {
        "exp": rng.uniform(-100, 88, 400000), "log": np.exp(rng.uniform(-80, 80, 400000)),
        "log1p": np.exp(rng.uniform(-16, 16, 400000)), "softplus": rng.uniform(-110, 110, 400000),
        "phi4": np.exp(rng.uniform(np.log(1e-8), np.log(30), 400000)),
        "phi2": np.exp(rng.uniform(np.log(1e-8), np.log(30), 400000)),
        "tanh": rng.uniform(-10, 10, 400000), "atanh": rng.uniform(-0.9999999, 0.9999999, 400000),
    }
    for fn, x in cases.items():
        x = x.astype(np.float32)
        dx = ctx.asarray(x)
        dy = ctx.empty(x.shape, np.float32)
        _ffi.call("fbgnn_math_probe", ctx.handle, fn.encode(), dx.ptr, dy.ptr, x.size)
        ref = oracle.math_fn({"exp": "expf", "log": "logf", "log1p": "log1pf_pos", "softplus": "softplusf",
                              "phi4": "phi4f", "phi2": "phi2f", "tanh": "tanhf", "atanh": "atanhf"}[fn], x)
        assert_bitexact(dy.numpy(), ref, f"math {fn}")


def test_dlpack_roundtrip():
    import fbgnn as F
    ctx = F.default_context()
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    d = ctx.asarray(a)
    v = d.transpose((1, 0, 2))
    assert np.array_equal(v.numpy(), a.transpose(1, 0, 2))
    e = F.from_dlpack(v)                                  # export + import, zero copy
    assert e.ptr == v.ptr and e.shape == v.shape and e.strides == v.strides
    assert np.array_equal(e.numpy(), a.transpose(1, 0, 2))


@pytest.mark.parametrize("large", ["cluster", "gstate"])
def test_code_larger_than_shared_memory(oracle, weights, large, monkeypatch):
    """[[7688,50]] hypergraph-product code: 36 bytes of decoder state per qubit = 277 KB per frame, more than an SM's
    227 KB of shared memory.  Default: a thread-block cluster decodes the frame with the messages in distributed shared
    memory (csrc/fbgnn_cluster.cuh); FBGNN_BP4_LARGE=gstate selects the fallback whose message arrays live in HBM / L2.
    Same arithmetic either way: layer outputs (incl. the final messages) and pipeline flags bit-exact with the oracle."""
    import fbgnn as F
    monkeypatch.setenv("FBGNN_BP4_LARGE", large)
    h = F.create_circulant_matrix(62, [0, 2, 5])
    code = F.hypergraph_product(h, h)
    assert (code.N, code.K) == (7688, 50)
    B = 5
    nx, nz, sx, sz = _noise_and_syndromes(oracle, code, B, 0.03, seed=5)
    g = oracle.CodeGraph(code)
    prior = oracle.prior_llr(0.03)
    rng = np.random.default_rng(1)
    llr = (prior + rng.normal(0, 0.3, (B, 3, code.N))).astype(np.float32)
    for cn_type, it in (("boxplus-phi", 6), ("minsum", 3)):
        dec = F.QLDPCBPDecoder(code, num_iter=it, normalization_factor=0.9, cn_type=cn_type, stage_one=True)
        out = dec((llr, sx, sz))
        ref = oracle.bp4(g, llr, sx, sz, it, 0.9, cn_type)
        for k, o in zip(("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit"), out):
            assert_bitexact(np.asarray(o, dtype=ref[k].dtype), ref[k], f"large code {cn_type} {k}")
        dev = dec._device()
        d = dec.decode_device(dev.ctx.asarray(llr), dev.ctx.asarray(sx), dev.ctx.asarray(sz), want_msgs=True)
        refm = oracle.bp4(g, llr, sx, sz, it, 0.9, cn_type, want_msgs=True)
        assert_bitexact(d[7].numpy(), refm["msg_x"], f"large code {cn_type} msg_x")
        assert_bitexact(d[8].numpy(), refm["msg_z"], f"large code {cn_type} msg_z")
    # constant prior (stage 0 of a pipeline) through the same path
    d1c = F.QLDPCBPDecoder(code, num_iter=5, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    outc = d1c.decode_device(None, d1c._device().ctx.asarray(sx), d1c._device().ctx.asarray(sz), prior=float(prior))
    refc = oracle.bp4(g, float(prior), sx, sz, 5, 1.0, "boxplus-phi")
    for k, o in zip(("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit"), outc):
        assert_bitexact(o.numpy(), refc[k], f"large code const prior {k}")
    G = F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True)
    G.set_weights(weights["c882"])
    d1 = F.QLDPCBPDecoder(code, num_iter=8, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=4, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d2], [G], num_layers=2, seed=3)
    res = model.run(6, 0.05, want_counters=True)
    ref = oracle.pipeline(g, [8, 4], [oracle.Gnn(weights["c882"])], 0.05, seed=3, B=6)
    assert_bitexact(res["flags"].numpy(), ref["flags"], "large code pipeline flags")
    assert res["counters"].tolist() == ref["counters"].tolist()
