"""Feedback GNN with its dense products on the tcgen05 tensor cores (``gemm="tf32x3"``, csrc/fbgnn_gnn_tc.cuh).  The
arithmetic of a tcgen05.mma kind::tf32 step is an integer model characterised on a B200 (csrc/fb_umma.h,
tests/test_umma_model.py), so the CPU oracle reproduces this form BIT FOR BIT too (oracle.Gnn(gemm="tf32x3")); against
the FMA form it agrees to float32 re-association accuracy (1e-5 of the largest output is the stated tolerance)."""
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
            g = oracle.CodeGraph(code)
            ref = oracle.gnn(g, oracle.Gnn(weights[name], "tanh", reduce_op), h_vn, lhx, lhz, sx, sz)
            nt = min(B, 40)                                                         # the emulation is slow: a slice
            ref_tc = oracle.gnn(g, oracle.Gnn(weights[name], "tanh", reduce_op, gemm="tf32x3"), h_vn[:nt],
                                np.ascontiguousarray(lhx[:, :nt]), np.ascontiguousarray(lhz[:, :nt]),
                                np.ascontiguousarray(sx[:, :nt]), np.ascontiguousarray(sz[:, :nt]))
    finally:
        ctx.set_math("exact")
    assert np.array_equal(outs["fma"].view(np.uint32), ref.view(np.uint32))          # the default stays bit-exact
    assert np.array_equal(outs["tf32x3"][:nt].view(np.uint32), ref_tc.view(np.uint32)), \
        f"tensor-core form differs from its oracle on {int(np.sum(outs['tf32x3'][:nt] != ref_tc))} values"
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


@pytest.mark.parametrize("arith,skip", [("exact", False), ("sfu", False), ("sfu", True)])
def test_pipeline_with_tensor_core_gnn_bitexact(codes, oracle, weights, arith, skip):
    """BP -> (GNN -> BP) x 2 with the tensor-core GNN against the oracle pipeline with the emulated tensor-core GNN:
    per-frame flags, residual errors and counters bit for bit."""
    import fbgnn as F
    code = codes["c882"]
    w = weights["c882"]
    ctx = F.default_context()
    ctx.set_math(arith)
    try:
        G = F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True, gemm="tf32x3")
        G.set_weights(w)
        d1 = F.QLDPCBPDecoder(code, num_iter=24, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        d2 = F.QLDPCBPDecoder(code, num_iter=8, normalization_factor=0.9, cn_type="boxplus-phi", stage_one=True)
        model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d2, d2], [G, G], num_layers=3, seed=7, first_frame=1000,
                                                   skip_inactive=skip)
        B, p = 128, 0.12
        res = model.run(B, p, want_counters=True)
        with oracle.math(arith):
            ref = oracle.pipeline(oracle.CodeGraph(code), [24, 8, 8], [oracle.Gnn(w, gemm="tf32x3")] * 2, p, p0=0.05,
                                  factors=[1.0, 0.9, 0.9], seed=7, first_frame=1000, B=B, skip_inactive=False, want_diff=True)
    finally:
        ctx.set_math("exact")
    assert np.array_equal(res["flags"].numpy(), ref["flags"])
    assert np.array_equal(res["x_diff"].numpy(), ref["x_diff"]) and np.array_equal(res["z_diff"].numpy(), ref["z_diff"])
    assert np.array_equal(res["counters"], ref["counters"])


def test_headline_pipeline_with_tensor_core_gnn_bitexact(c1270, oracle, weights):
    """The bench configuration -- [[1270,28]], BP64 -> (GNN -> BP16) x 3, p = 0.10, SFU arithmetic, tensor-core GNN --
    frame by frame against the oracle (MUFU tables + emulated tensor-core steps)."""
    import fbgnn as F
    w = weights["c1270"]
    ctx = F.default_context()
    ctx.set_math("sfu")
    try:
        G = F.Feedback_GNN(c1270, 20, 40, 2, "mean", "tanh", True, gemm="tf32x3")
        G.set_weights(w)
        d1 = F.QLDPCBPDecoder(c1270, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        d2 = F.QLDPCBPDecoder(c1270, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        model = F.Sandwich_BP_GNN_Evaluation_Model(c1270, [d1, d2, d2, d2], [G] * 3, num_layers=4, seed=2, first_frame=5000)
        B, p = 96, 0.10
        res = model.run(B, p, want_counters=True)
        with oracle.math("sfu"):
            ref = oracle.pipeline(oracle.CodeGraph(c1270), [64, 16, 16, 16], [oracle.Gnn(w, gemm="tf32x3")] * 3, p, p0=0.05,
                                  seed=2, first_frame=5000, B=B, skip_inactive=False, want_diff=True)
    finally:
        ctx.set_math("exact")
    assert np.array_equal(res["flags"].numpy(), ref["flags"])
    assert np.array_equal(res["x_diff"].numpy(), ref["x_diff"]) and np.array_equal(res["z_diff"].numpy(), ref["z_diff"])
    assert np.array_equal(res["counters"], ref["counters"])
    assert int(ref["counters"][3]) > 10                      # frames did reach the GNN rounds


def test_this_gpus_tensor_core_is_the_model(oracle):
    """One tcgen05.mma kind::tf32 step on fresh random operands (TF32-exact and full 24-bit significands, exponents over
    2^-14 .. 2^14, with and without accumulator, sparse rows): the GPU the tests run on returns, on every one of the
    262 144 outputs, the bits of the integer model the repository carries (csrc/fb_umma.h)."""
    import ctypes as C
    import fbgnn as F
    from fbgnn import _ffi
    ctx = F.default_context()
    rng = np.random.default_rng(2026)
    T = 128

    def rnd(shape, bits, span):
        e = rng.integers(-span, span + 1, shape)
        m = (1 << (bits - 1)) + rng.integers(0, 1 << (bits - 1), shape)
        v = np.ldexp(m.astype(np.float64), e - (bits - 1)) * rng.choice([-1.0, 1.0], shape)
        return v.astype(np.float32)

    A = np.empty((T, 128, 8), np.float32); B = np.empty((T, 8, 16), np.float32); Din = np.zeros((T, 128, 16), np.float32)
    for tr in range(T):
        span = 14 if tr & 1 else 2
        bits = 24 if (tr >> 2) & 1 else 11
        A[tr], B[tr] = rnd((128, 8), bits, span), rnd((8, 16), bits, span)
        if tr & 2:
            Din[tr] = rnd((128, 16), 24, 2 * span)
        if tr % 16 == 9:
            A[tr][rng.random((128, 8)) < 0.3] = 0.0
    dA, dB, dD = ctx.asarray(A), ctx.asarray(B), ctx.asarray(Din)
    dO = ctx.empty((T, 128, 16), np.float32)
    _ffi.call("fbgnn_umma_probe", ctx.handle, dA.ptr, dB.ptr, dD.ptr, dO.ptr, T)
    out = np.ascontiguousarray(dO.numpy())
    L = oracle.lib()
    L.orc_umma8_check.restype = C.c_int64
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.orc_umma8_check(p(A), p(B), p(Din), p(out), C.c_int32(T)) == 0
