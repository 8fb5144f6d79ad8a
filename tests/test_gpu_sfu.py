"""The SFU arithmetic (Context.set_math("sfu"); csrc/fb_math.h): exp / log on MUFU.EX2 / MUFU.LG2 with range
reductions that confine the MUFU inputs to finite sets, which the CPU oracle evaluates through tables measured on
a B200 (tests/golden/sfu_b200_*.xz).  Same bar as the exact arithmetic: every float32 output BIT-EXACT against the
oracle in the same arithmetic, and the reference's published logical error rates reproduced.  The parity cases are
the ones of test_gpu_parity.py / test_gpu_baseline_configs.py, re-run with both sides switched to SFU arithmetic."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def sfu(oracle):
    import fbgnn as F
    ctx = F.default_context()
    ctx.set_math("sfu")
    oracle.set_math("sfu")
    assert ctx.get_math() == "sfu" and oracle.get_math() == "sfu"
    yield ctx
    ctx.set_math("exact")
    oracle.set_math("exact")


def test_hardware_matches_the_committed_tables(sfu, oracle):
    """The MUFU of this GPU returns, on every table entry, the bits the repository carries (so the oracle's SFU
    arithmetic is this GPU's), and the composed functions agree bit for bit on dense samples."""
    from fbgnn import _ffi
    from oracle import sfu_tables as T
    import test_gpu_parity as P
    ctx = sfu

    def probe(fn, x):
        dx = ctx.asarray(x)
        dy = ctx.empty(x.shape, np.float32)
        _ffi.call("fbgnn_math_probe", ctx.handle, fn.encode(), dx.ptr, dy.ptr, x.size)
        return dy.numpy()

    ex2, lg2, lg2b, rcp = T.tables()
    rng = np.random.default_rng(1)
    for fn, inputs, table in (("mufu_ex2", T.ex2_inputs(), ex2), ("mufu_lg2", T.lg2_inputs(), lg2),
                              ("mufu_lg2", T.lg2b_inputs()[1:], lg2b[1:]), ("mufu_rcp", T.rcp_inputs(), rcp)):
        P.assert_bitexact(probe(fn, inputs), table, fn)             # exhaustive: every table entry
    cases = {"sfu_exp": ("sfu_expf", rng.uniform(-100, 88, 400000)),
             "sfu_log": ("sfu_logf", np.exp(rng.uniform(-80, 80, 400000))),
             "sfu_softplus": ("m_softplusf", rng.uniform(-110, 110, 400000)),
             "sfu_phi4": ("m_phi4f", np.exp(rng.uniform(np.log(1e-8), np.log(30), 400000))),
             "sfu_phi2": ("m_phi2f", np.exp(rng.uniform(np.log(1e-8), np.log(30), 400000))),
             "sfu_tanh": ("m_tanhf", rng.uniform(-20, 20, 400000)),
             "sfu_atanh": ("m_atanhf", rng.uniform(-0.9999999, 0.9999999, 400000))}
    for fn, (ofn, x) in cases.items():
        x = x.astype(np.float32)
        P.assert_bitexact(probe(fn, x), oracle.math_fn(ofn, x), fn)


@pytest.mark.parametrize("name,B,p", [("steane", 64, 0.05), ("rsurf3", 64, 0.08), ("gb48", 64, 0.05), ("c882", 96, 0.09)])
@pytest.mark.parametrize("cn_type,factor", [("boxplus-phi", 1.0), ("boxplus-phi", 0.625), ("minsum", 0.8), ("boxplus", 1.0)])
def test_bp4_layer_bitexact_sfu(sfu, codes, oracle, name, B, p, cn_type, factor):
    import test_gpu_parity as P
    P.test_bp4_layer_bitexact(codes, oracle, name, B, p, cn_type, factor)


@pytest.mark.parametrize("reduce_op", ["mean", "max"])
def test_gnn_layers_bitexact_sfu(sfu, codes, c1270, oracle, weights, reduce_op):
    """Feedback_GNN and GNN_BP4 with tanh on the SFU (exp + MUFU.RCP)."""
    import test_gpu_baseline_configs as C
    import test_gpu_parity as P
    import test_gnn_bp4 as G
    P.test_gnn_layer_bitexact(codes, oracle, weights, "c882", "c882", reduce_op)
    if reduce_op == "mean":
        C.test_gnn_layer_bitexact_c1270(c1270, oracle, weights)
        P.test_gnn_irregular_code_and_relu(codes, oracle)
        G.test_cuda_gnn_bp4_bitexact(oracle, codes, "c882", "mean", True)


@pytest.mark.parametrize("name", ["rsurf3", "c882"])
def test_bp2_and_stage_two_bitexact_sfu(sfu, codes, oracle, name):
    import test_gpu_parity as P
    P.test_bp2_layer_bitexact(codes, oracle, name, "boxplus-phi")
    if name == "c882":
        P.test_bp4_stage_two_per_iteration_soft_syndromes(codes, oracle, name)


@pytest.mark.parametrize("it", [1, 16, 64])
@pytest.mark.parametrize("prior_kind", ["const", "per_variable"])
def test_bp4_layer_bitexact_c1270_sfu(sfu, c1270, oracle, it, prior_kind):
    import test_gpu_baseline_configs as C
    C.test_bp4_layer_bitexact_c1270(c1270, oracle, it, prior_kind)


@pytest.mark.parametrize("skip", [False, True])
def test_pipelines_bitexact_sfu(sfu, codes, c1270, oracle, weights, skip):
    import test_gpu_baseline_configs as C
    import test_gpu_parity as P
    C.test_pipeline_bitexact_c1270_three_rounds(c1270, oracle, weights, 0.10, skip)
    P.test_pipeline_bitexact(codes, oracle, weights, "c882", 3, 0.12, 160, skip)
    P.test_pipeline_bitexact(codes, oracle, weights, "rsurf3", 0, 0.08, 300, skip)


def test_config0_fixed_point_exit_bsc_sfu(sfu, codes, c1270, oracle):
    import test_gpu_baseline_configs as C
    import test_gpu_parity as P
    C.test_config0_literally(codes, oracle, 0.05)
    C.test_fixed_point_exit_is_exact(codes, c1270, oracle, "c1270", 0.02)
    C.test_fixed_point_exit_is_exact(codes, c1270, oracle, "c882", 0.02)
    P.test_bsc_pipeline_bitexact(codes, oracle, "c882", True)
    C.test_bp2_layer_bitexact_c1270(c1270, oracle)


def _model(F, code, wfile, nG, **kw):
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)
    F.load_weights(G, os.path.join(F.WEIGHTS_DIR, wfile))
    d1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    return F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [G] * nG, num_layers=nG + 1, **kw)


def _compatible(k, n, k_pub, n_pub, z=3.7):
    p_pool = (k + k_pub) / (n + n_pub)
    sigma = np.sqrt(p_pool * (1 - p_pool) * (1 / n + 1 / n_pub))
    return abs(k / n - k_pub / n_pub) < z * sigma + 1e-12


@pytest.mark.parametrize("p,k_pub,n_pub,frames", [(0.14, 1986, 5000, 20000), (0.13, 705, 5000, 20000),
                                                  (0.12, 139, 5000, 40000), (0.11, 106, 25000, 100000),
                                                  (0.10, 100, 275000, 600000)])
def test_published_ler_c1270_three_rounds_sfu(sfu, c1270, p, k_pub, n_pub, frames):
    """examples/n1270.ipynb cell 2 in SFU arithmetic: the block error counts must be statistically compatible with
    the published ones (two-sample binomial, as for the exact arithmetic in test_gpu_fullsize.py)."""
    import fbgnn as F
    model = _model(F, c1270, "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy", 3, seed=1000 + int(p * 100),
                   skip_inactive=True)
    k = 0
    for _ in range(frames // 20000):
        k += int(model.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"][2])
    assert _compatible(k, frames, k_pub, n_pub), (p, k, frames, k_pub, n_pub)


@pytest.mark.parametrize("p,bler_pub,frames", [(0.12, 6.74e-2, 20000), (0.11, 1.52e-2, 40000), (0.10, 2.40e-3, 200000)])
def test_published_ler_c882_five_rounds_sfu(sfu, codes, p, bler_pub, frames):
    """examples/n882.ipynb cell 3 (nG = 5) in SFU arithmetic."""
    import fbgnn as F
    model = _model(F, codes["c882"], "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy", 5, seed=2000 + int(p * 100),
                   skip_inactive=True)
    k = 0
    for _ in range(frames // 20000):
        k += int(model.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"][2])
    n_pub = max(int(round(100 / bler_pub)), 5000)
    assert _compatible(k, frames, int(round(bler_pub * n_pub)), n_pub), (p, k, frames)


def test_sfu_and_exact_agree_where_bp_is_not_chaotic(codes):
    """Same frames through both arithmetics at low noise: the hard decisions coincide on almost every frame and the
    logical outcome on all but a few (at p = 0.03 even a re-ordered sum flips ~0.5 % of the frames)."""
    import fbgnn as F
    ctx = F.default_context()
    code = codes["c882"]
    B, p = 20000, 0.03
    W = "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"
    ctx.set_math("exact")
    a = _model(F, code, W, 1, seed=77).run(B, p)
    fa, xa = a["flags"].numpy(), a["x_diff"].numpy()
    ctx.set_math("sfu")
    try:
        b = _model(F, code, W, 1, seed=77).run(B, p)
        fb, xb = b["flags"].numpy(), b["x_diff"].numpy()
    finally:
        ctx.set_math("exact")
    assert np.all(xa == xb, axis=1).mean() > 0.995
    assert np.mean((fa & 3) == (fb & 3)) > 0.9995
