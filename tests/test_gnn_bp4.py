"""GNN_BP4 (gnn.py:71-751, BASELINE configs[4]).  The reference ships neither weights nor recorded
outputs for this layer and its call() is broken at HEAD, so parity here is UNPINNED:
the C oracle is cross-checked against the independent numpy restatement (CPU), and the CUDA path
is compared bit for bit with the C oracle (GPU)."""
import numpy as np
import pytest


def _weights_list(W, use_bias=True):
    from oracle.c_oracle import GBP_KEYS
    out = [W["Winv"], W["binv"]]
    for k in GBP_KEYS:
        out += list(W[k])
    return out if use_bias else out[0::2]


def _syndromes(oracle, code, B, p, seed):
    nx, nz = oracle.pauli(seed, 0, B, code.N, p)
    sx = ((code.hx @ nz.T.astype(np.int64)) & 1).T.astype(np.uint8)
    sz = ((code.hz @ nx.T.astype(np.int64)) & 1).T.astype(np.uint8)
    return np.ascontiguousarray(sx), np.ascontiguousarray(sz)


@pytest.mark.parametrize("name,reduce_op", [("c882", "mean"), ("rsurf3", "sum"), ("toric4", "max")])
def test_c_oracle_matches_numpy_restatement(oracle, codes, name, reduce_op):
    from oracle import np_oracle as N
    code = codes[name]
    g = oracle.CodeGraph(code)
    W = oracle.gnn_bp4_random_weights(seed=4)
    sx, sz = _syndromes(oracle, code, 5, 0.06, 3)
    r = oracle.gnn_bp4(g, W, sx, sz, 3, reduce_op=reduce_op)
    out, xh, zh = N.gnn_bp4(N.Side(code.hx), N.Side(code.hz), N.Side(code.lx), N.Side(code.lz), W, sx, sz, 3,
                            reduce_op=reduce_op)
    for it in range(3):
        assert np.allclose(out[it][0], r["x_logit"][it], rtol=2e-3, atol=1e-5)
        assert np.allclose(out[it][1], r["z_logit"][it], rtol=2e-3, atol=1e-5)
    assert r["x_logit"].shape == (3, code.hz.shape[0] + code.lz.shape[0], 5)
    assert np.mean(xh == r["x_hat"]) > 0.999 and np.mean(zh == r["z_hat"]) > 0.999


def test_layer_weight_contract(codes):
    import fbgnn as F
    G = F.GNN_BP4(codes["steane"], num_embed_dims=20, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2,
                  num_iter=4, use_bias=True)
    w = G.get_weights()
    assert len(w) == 30 and w[0].shape == (20, 3) and np.all(w[0] == 0) and np.all(w[1] == 1)
    assert w[2].shape == (40, 40) and w[10].shape == (41, 40) and w[26].shape == (60, 40) and w[28].shape == (40, 20)
    assert sum(a.size for a in w) == 2 * 20 * 3 // 2 + 3 + 4 * 2460 + 2 * 2500 + 3260
    assert len(F.GNN_BP4(codes["steane"], 20, 20, 40, 2, 4, use_bias=False).get_weights()) == 15
    with pytest.raises(ValueError):
        G.set_weights(w[:29])
    with pytest.raises(NotImplementedError):
        F.GNN_BP4(codes["steane"], 20, 20, 40, 3, 4)
    with pytest.raises(NotImplementedError):
        F.GNN_BP4(codes["steane"], 20, 20, 40, 2, 4, use_attributes=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,reduce_op,bias", [("c882", "mean", True), ("rsurf3", "sum", True), ("toric4", "min", False),
                                                  ("gb48", "mean", True)])
def test_cuda_gnn_bp4_bitexact(oracle, codes, name, reduce_op, bias):
    import fbgnn as F
    code = codes[name]
    W = oracle.gnn_bp4_random_weights(seed=4, use_bias=bias)
    B, it = 37, 4
    sx, sz = _syndromes(oracle, code, B, 0.06, 9)
    G = F.GNN_BP4(code, num_embed_dims=20, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, num_iter=it,
                  reduce_op=reduce_op, activation="tanh", use_bias=bias)
    G.set_weights(_weights_list(W, bias))
    llr_hat, x_hat, z_hat = G((sx, sz))
    ref = oracle.gnn_bp4(oracle.CodeGraph(code), W, sx, sz, it, reduce_op=reduce_op)
    assert len(llr_hat) == it and x_hat.shape == (code.N, B)
    for i in range(it):
        for got, want in ((llr_hat[i][0], ref["x_logit"][i]), (llr_hat[i][1], ref["z_logit"][i])):
            assert got.shape == want.shape
            assert np.array_equal(np.ascontiguousarray(got).view(np.uint32), np.ascontiguousarray(want).view(np.uint32)), (name, i)
    assert np.array_equal(x_hat.astype(np.uint8), ref["x_hat"]) and np.array_equal(z_hat.astype(np.uint8), ref["z_hat"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,reduce_op,bias", [("c882", "mean", True), ("rsurf3", "sum", True), ("gb48", "mean", False)])
def test_cuda_gnn_bp4_tensor_core_path(oracle, codes, name, reduce_op, bias):
    """gemm="tf32x3": the per-node matrix products on tcgen05 tensor cores with the three-product TF32 split.
    Not bit-exact by construction (the tensor core's accumulation order is its own); stated tolerance against
    the oracle: |logit difference| <= 2e-5 + 1e-5 |logit| in every iteration, identical hard decisions on
    >= 99.99 % of the qubits."""
    import fbgnn as F
    code = codes[name]
    W = oracle.gnn_bp4_random_weights(seed=4, use_bias=bias)
    B, it = 300, 6                         # 300 frames: several 128-row tiles with a ragged tail
    sx, sz = _syndromes(oracle, code, B, 0.06, 9)
    G = F.GNN_BP4(code, num_embed_dims=20, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, num_iter=it,
                  reduce_op=reduce_op, activation="tanh", use_bias=bias, gemm="tf32x3")
    G.set_weights(_weights_list(W, bias))
    llr_hat, x_hat, z_hat = G((sx, sz))
    ref = oracle.gnn_bp4(oracle.CodeGraph(code), W, sx, sz, it, reduce_op=reduce_op)
    for i in range(it):
        for got, want in ((llr_hat[i][0], ref["x_logit"][i]), (llr_hat[i][1], ref["z_logit"][i])):
            assert got.shape == want.shape
            assert np.all(np.abs(got - want) <= 2e-5 + 1e-5 * np.abs(want)), (name, i, float(np.abs(got - want).max()))
    agree = np.mean((x_hat.astype(np.uint8) == ref["x_hat"]) & (z_hat.astype(np.uint8) == ref["z_hat"]))
    assert agree >= 0.9999, agree


@pytest.mark.gpu
def test_tensor_core_path_rejects_unsupported_configurations(codes):
    import fbgnn as F
    G = F.GNN_BP4(codes["steane"], 20, 20, 40, 2, 2, reduce_op="max", gemm="tf32x3")
    sx = np.zeros((2, codes["steane"].hx.shape[0]), np.uint8)
    with pytest.raises(F.FbgnnError):
        G((sx, sx))
    with pytest.raises(ValueError):
        F.GNN_BP4(codes["steane"], 20, 20, 40, 2, 2, gemm="bf16")
