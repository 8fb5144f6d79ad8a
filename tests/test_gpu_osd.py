"""BP + OSD-0 (bp_osd.py): the CUDA path against the oracle, and the reference's published
BP+OSD-0 error rates (examples/OSD.ipynb cells 2, 3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_osd0_layer_bitexact_and_solves_the_syndrome(codes, oracle):
    import fbgnn as F
    for name in ("c882", "gb48", "rsurf3"):
        code = codes[name]
        n = code.N
        rng = np.random.default_rng(1)
        B = 40
        noise = (rng.random((B, n)) < 0.06).astype(np.uint8)
        synd = ((code.hx @ noise.T.astype(np.int64)) & 1).astype(np.uint8)
        red = synd[code.pivot_hx]
        llr = rng.normal(3, 1, (B, n)).astype(np.float32)
        llr[noise == 1] -= 2.5
        llr[:, ::7] = 1.25                                  # ties: broken by index on both sides
        basis = code.hx[code.pivot_hx]
        e = F.OSD0_Decoder(n)(llr, np.broadcast_to(basis, (B,) + basis.shape), red, B)
        ref = oracle.osd0(basis, llr, red)
        assert e.dtype == bool and np.array_equal(e.astype(np.uint8), ref)
        assert np.array_equal((code.hx @ e.T.astype(np.int64)) & 1, synd)   # satisfies the full syndrome


@pytest.mark.parametrize("skip", [False, True])
def test_bp4_osd_pipeline_bitexact(codes, oracle, skip):
    import fbgnn as F
    code = codes["c882"]
    dec = F.QLDPCBPDecoder(code, num_iter=30, normalization_factor=0.8, cn_type="minsum", stage_one=True)
    model = F.BP4_OSD_Model(code, dec, F.OSD0_Decoder(code.N), seed=4)
    model._inner.skip_inactive = skip
    B, p = 600, 0.11
    res = model.run(B, p, want_counters=True)
    ref = oracle.pipeline(oracle.CodeGraph(code), [30], [], p, p0=None, factors=[0.8], cn_types=["minsum"], seed=4,
                          B=B, osd0=True, want_diff=True)
    assert np.array_equal(res["flags"].numpy(), ref["flags"])
    assert np.array_equal(res["x_diff"].numpy(), ref["x_diff"]) and np.array_equal(res["z_diff"].numpy(), ref["z_diff"])
    assert np.array_equal(res["counters"], ref["counters"])
    assert res["counters"][1] == 0 and res["counters"][3] == 0      # OSD always meets the syndrome
    s0, ls = model(B, p)
    assert s0.count_nonzero_rows() == 0 and ls.shape == (B, 2 * code.K)


def test_sandwich_then_osd_bitexact(codes, oracle, weights):
    """OSD-0 after BP -> GNN -> BP (not a reference model, but the same building blocks)."""
    import fbgnn as F
    code = codes["c882"]
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    G.set_weights(weights["c882"])
    d = F.QLDPCBPDecoder(code, num_iter=12, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d, d], [G], num_layers=2, seed=8, osd0=True)
    res = model.run(300, 0.12, want_counters=True)
    ref = oracle.pipeline(oracle.CodeGraph(code), [12, 12], [oracle.Gnn(weights["c882"])], 0.12, seed=8, B=300, osd0=True)
    assert np.array_equal(res["flags"].numpy(), ref["flags"]) and np.array_equal(res["counters"], ref["counters"])


def test_bp2_osd_pipeline_bitexact(codes, oracle):
    import fbgnn as F
    code = codes["c882"]
    dec = F.LDPCBPDecoder(code.hx, is_syndrome=True, hard_out=False, cn_type="minsum", num_iter=30, normalization_factor=0.8)
    model = F.BP2_OSD_Model(code.hx, code.hx_basis, code.pivot_hx, code.lx, dec, F.OSD0_Decoder(code.N), seed=6)
    B, p = 3000, 0.06
    res = model.run(B, p, want_counters=True)
    ref = oracle.bsc_pipeline(code.hx, code.lx, 30, p, factor=0.8, cn_type="minsum", seed=6, B=B,
                              osd_basis=code.hx_basis, osd_pivot=code.pivot_hx)
    assert np.array_equal(res["flags"].numpy(), ref["flags"]) and np.array_equal(res["counters"], ref["counters"])
    assert res["counters"][1] == 0


def _compatible(k, n, k_pub, n_pub, z=3.7):
    p_pool = (k + k_pub) / (n + n_pub)
    sigma = np.sqrt(p_pool * (1 - p_pool) * (1 / n + 1 / n_pub))
    return abs(k / n - k_pub / n_pub) < z * sigma + 1e-12


def test_published_bp_osd_error_rates(codes):
    """examples/OSD.ipynb cell 2: BP4(minsum, 100 it., f=0.8)+OSD0 on [[882,24]]: 111/300000 at p=0.10,
    102/1700000 at p=0.09; cell 3: BP2+OSD0 on hx over the BSC: 117/200000 at p=0.05."""
    import fbgnn as F
    code = codes["c882"]
    dec = F.QLDPCBPDecoder(code, num_iter=100, normalization_factor=0.8, cn_type="minsum", stage_one=True)
    model = F.BP4_OSD_Model(code, dec, F.OSD0_Decoder(code.N), seed=70)
    model._inner.skip_inactive = True
    for p, k_pub, n_pub, frames in ((0.10, 111, 300000, 300000), (0.09, 102, 1700000, 600000)):
        k = sum(int(model.run(50000, p, want_flags=False, want_diff=False, want_counters=True)["counters"][2])
                for _ in range(frames // 50000))
        assert _compatible(k, frames, k_pub, n_pub), (p, k, frames)
    dec2 = F.LDPCBPDecoder(code.hx, is_syndrome=True, hard_out=False, cn_type="minsum", num_iter=100, normalization_factor=0.8)
    m2 = F.BP2_OSD_Model(code.hx, code.hx_basis, code.pivot_hx, code.lx, dec2, F.OSD0_Decoder(code.N), seed=71)
    k = sum(int(m2.run(50000, 0.05, want_counters=True)["counters"][2]) for _ in range(4))
    assert _compatible(k, 200000, 117, 200000), k
