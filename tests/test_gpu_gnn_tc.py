"""Feedback GNN with its dense products on the tcgen05 tensor cores (``gemm="tf32x3"``, csrc/fbgnn_gnn_tc.cuh): an opt-in
form that agrees with the default FMA kernel / the oracle to float32 re-association accuracy, not bit for bit.  The
tolerance is stated here: 1e-5 of the largest output per call (north star: FP32 quantities within 1e-5 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(oracle, code, B, p, seed, iters=8):
    nx, nz = oracle.pauli(seed, 0, B, code.N, p)
    sx = ((code.hx @ nz.T.astype(np.int64)) & 1).astype(np.uint8)
    sz = ((code.hz @ nx.T.astype(np.int64)) & 1).astype(np.uint8)
    r = oracle.bp4(oracle.CodeGraph(code), float(oracle.prior_llr(0.05)), sx, sz, iters)
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1).astype(np.float32)
    return h_vn, r["z_logit"], r["x_logit"], sx, sz


@pytest.mark.parametrize("arith", ["exact", "sfu"])
@pytest.mark.parametrize("name,reduce_op,B", [("c882", "mean", 300), ("c882", "sum", 37), ("c1270", "mean", 130)])
def test_gnn_tensor_core_form_within_tolerance(codes, c1270, oracle, weights, name, reduce_op, B, arith):
    import fbgnn as F
    code = c1270 if name == "c1270" else codes[name]
    h_vn, lhx, lhz, sx, sz = _inputs(oracle, code, B, 0.09, seed=21)
    ctx = F.default_context()
    ctx.set_math(arith)
    try:
        outs = {}
        for gemm in ("fma", "tf32x3"):
            G = F.Feedback_GNN(code, 20, 40, 2, reduce_op, "tanh", True, gemm=gemm)
            G.set_weights(weights[name])
            outs[gemm] = np.asarray(G((h_vn, lhx, lhz, sx, sz)))
        with oracle.math(arith):
            ref = oracle.gnn(oracle.CodeGraph(code), oracle.Gnn(weights[name], "tanh", reduce_op), h_vn, lhx, lhz, sx, sz)
    finally:
        ctx.set_math("exact")
    assert np.array_equal(outs["fma"].view(np.uint32), ref.view(np.uint32))          # the default stays bit-exact
    scale = float(np.abs(ref).max())
    err = float(np.abs(outs["tf32x3"] - ref).max())
    assert err <= 1e-5 * scale, (err, scale)
    assert not np.array_equal(outs["tf32x3"], ref) or B < 8                        # it really is a different evaluation


def test_gnn_tensor_core_form_rejects_other_configurations(codes, weights):
    import fbgnn as F
    code = codes["c882"]
    G = F.Feedback_GNN(code, 20, 40, 2, "max", "tanh", True, gemm="tf32x3")
    G.set_weights(weights["c882"])
    with pytest.raises(F.FbgnnError):
        G.device_handle()
    with pytest.raises(ValueError):
        F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True, gemm="fp8")
    irregular = codes["rsurf3"]                                                     # not (3,.)-regular: refused at launch
    G = F.Feedback_GNN(irregular, 20, 40, 2, "mean", "tanh", True, gemm="tf32x3")
    G.set_weights(weights["c882"])
    n, mx, mz = irregular.N, irregular.hx.shape[0], irregular.hz.shape[0]
    with pytest.raises(F.FbgnnError):
        G((np.zeros((2, n, 3), np.float32), np.zeros((mx, 2), np.float32), np.zeros((mz, 2), np.float32),
           np.zeros((mx, 2), np.uint8), np.zeros((mz, 2), np.uint8)))


def test_pipeline_with_tensor_core_gnn_tracks_the_bit_exact_pipeline(codes, oracle, weights):
    """BP -> (GNN -> BP) x 2 on [[882,24]]: the per-frame indicators of the tensor-core form agree with the bit-exact
    pipeline on nearly every frame (differences are the float32 noise floor of a chaotic decoder) and the
    block-error counts agree within binomial noise."""
    import fbgnn as F
    code = codes["c882"]
    res = {}
    for gemm in ("fma", "tf32x3"):
        G = F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True, gemm=gemm)
        G.set_weights(weights["c882"])
        d1 = F.QLDPCBPDecoder(code, num_iter=32, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d2, d2], [G, G], num_layers=3, seed=77)
        r = model.run(20000, 0.10, want_counters=True)
        res[gemm] = (r["flags"].numpy(), r["counters"].tolist())
    agree = float(np.mean(res["fma"][0] == res["tf32x3"][0]))
    assert agree > 0.985, agree
    ka, kb = res["fma"][1][2], res["tf32x3"][1][2]
    assert abs(ka - kb) <= 4.0 * np.sqrt(max(ka, kb, 1)) + 5, (ka, kb)
