"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU: the oracle must still reproduce them bit for bit.  GPU: the CUDA path must, too."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                                                     np.ascontiguousarray(b, np.float32).view(np.uint32))
    return a.shape == b.shape and np.array_equal(a, b)


# ------------------------------------------------------------------ CPU: oracle vs golden
def test_oracle_reproduces_bp4_and_gnn_golden(oracle, codes, weights):
    code = codes["c882"]
    g = oracle.CodeGraph(code)
    d = load("bp4_c882_cfg0.npz")
    nx, nz = oracle.pauli(100, 0, 8, 882, 0.09)
    assert same(nx, d["noise_x"]) and same(nz, d["noise_z"])
    assert same((code.hx @ nz.T.astype(np.int64)) & 1, d["syndrome_x"])
    r = oracle.bp4(g, float(d["prior"]), d["syndrome_x"], d["syndrome_z"], 32, 0.625, "boxplus-phi")
    for k in ("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit"):
        assert same(r[k], d[k]), k
    e = load("gnn_c882.npz")
    out = oracle.gnn(g, oracle.Gnn(weights["c882"]), e["h_vn"], e["logit_hx"], e["logit_hz"], e["syndrome_x"],
                     e["syndrome_z"])
    assert same(out, e["out"])


def test_oracle_reproduces_bp2_and_pipeline_golden(oracle, codes, c1270, weights):
    d = load("bp2_c1270_cfg1.npz")
    soft, hard = oracle.bp2(c1270.hx, d["llr"], d["syndrome"], 64)
    assert same(soft, d["soft"]) and same(hard, d["hard"])
    e = load("pipeline_c882_nG5.npz")
    G = oracle.Gnn(weights["c882"])
    r = oracle.pipeline(oracle.CodeGraph(codes["c882"]), [64] + [16] * 5, [G] * 5, float(e["p"]), seed=int(e["seed"]),
                        first_frame=int(e["first_frame"]), B=len(e["flags"]), skip_inactive=True)
    assert same(r["flags"], e["flags"]) and same(r["counters"], e["counters"])


# ------------------------------------------------------------------ GPU: CUDA vs golden
@pytest.mark.gpu
def test_cuda_reproduces_layer_goldens(codes, c1270, weights):
    import fbgnn as F
    code = codes["c882"]
    d = load("bp4_c882_cfg0.npz")
    B = d["Lx"].shape[0]
    dec = F.QLDPCBPDecoder(code, num_iter=32, normalization_factor=0.625, cn_type="boxplus-phi", stage_one=True)
    llr = np.full((B, 3, code.N), d["prior"], np.float32)
    out = dec((llr, d["syndrome_x"], d["syndrome_z"]))
    for k, o in zip(("Lx", "Ly", "Lz", "x_hat", "z_hat", "x_logit", "z_logit"), out):
        assert same(np.asarray(o, d[k].dtype), d[k]), k
    e = load("gnn_c882.npz")
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    G.set_weights(weights["c882"])
    assert same(G((e["h_vn"], e["logit_hx"], e["logit_hz"], e["syndrome_x"], e["syndrome_z"])), e["out"])
    b = load("bp2_c1270_cfg1.npz")
    soft = F.LDPCBPDecoder(c1270.hx, is_syndrome=True, num_iter=64, hard_out=False)((b["llr"], b["syndrome"]))
    hard = F.LDPCBPDecoder(c1270.hx, is_syndrome=True, num_iter=64)((b["llr"], b["syndrome"]))
    assert same(soft, b["soft"]) and same(hard.astype(np.uint8), b["hard"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,nG,wkey", [("pipeline_c1270_nG1.npz", 1, "c1270"), ("pipeline_c882_nG5.npz", 5, "c882")])
def test_cuda_reproduces_pipeline_goldens(codes, c1270, weights, name, nG, wkey):
    import fbgnn as F
    code = c1270 if wkey == "c1270" else codes["c882"]
    e = load(name)
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    G.set_weights(weights[wkey])
    d1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    for skip in (False, True):
        model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [G] * nG, num_layers=nG + 1,
                                                   seed=int(e["seed"]), first_frame=int(e["first_frame"]),
                                                   skip_inactive=skip)
        res = model.run(len(e["flags"]), float(e["p"]), want_counters=True)
        assert same(res["flags"].numpy(), e["flags"]) and same(res["counters"], e["counters"])
