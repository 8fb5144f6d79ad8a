"""csrc/fb_umma.h: the integer model of one tcgen05.mma kind::tf32 step, replayed against what a B200 returned
(tests/golden/umma_b200_probe.npz: a slice of the dump of tools/micro/umma_probe.cu -- D_out = A B + D_in for K = 8, operands
with TF32-exact and with full 24-bit significands, exponents spread over 2^-14 .. 2^14, with and without accumulator)."""
import ctypes as C
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_umma_step_model_reproduces_the_hardware_bit_for_bit(oracle):
    Z = np.load(os.path.join(GOLDEN, "umma_b200_probe.npz"))
    A, B, Din, Dout = (np.ascontiguousarray(Z[k], np.float32) for k in ("A", "B", "Din", "Dout"))
    T = A.shape[0]
    assert A.shape == (T, 128, 8) and B.shape == (T, 8, 16) and Din.shape == Dout.shape == (T, 128, 16)
    L = oracle.lib()
    L.orc_umma8_check.restype = C.c_int64
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.orc_umma8_check(p(A), p(B), p(Din), p(Dout), C.c_int32(T)) == 0
    # the probe is not trivial: float32 FMAs in order give other bits for a large share of the outputs
    acc = Din.copy()
    for k in range(8):
        acc = (acc.astype(np.float64) + A[:, :, k, None].astype(np.float64) * B[:, None, k, :].astype(np.float64)).astype(np.float32)
    assert np.mean(acc.view(np.uint32) != Dout.view(np.uint32)) > 0.2


def test_tensor_core_gnn_oracle_tracks_the_fma_oracle(oracle, codes, weights):
    """The tensor-core form of the feedback GNN (three-product TF32 split, emulated exactly) stays within float32
    re-association accuracy of the FMA form."""
    code = codes["c882"]
    B = 3
    nx, nz = oracle.pauli(21, 0, B, code.N, 0.09)
    sx = ((code.hx @ nz.T.astype(np.int64)) & 1).astype(np.uint8)
    sz = ((code.hz @ nx.T.astype(np.int64)) & 1).astype(np.uint8)
    g = oracle.CodeGraph(code)
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 8)
    h = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1).astype(np.float32)
    for red in ("mean", "sum"):
        a = oracle.gnn(g, oracle.Gnn(weights["c882"], "tanh", red), h, r["z_logit"], r["x_logit"], sx, sz)
        b = oracle.gnn(g, oracle.Gnn(weights["c882"], "tanh", red, gemm="tf32x3"), h, r["z_logit"], r["x_logit"], sx, sz)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max() and not np.array_equal(a, b)
