"""Generates the committed golden vectors under tests/golden/ from the CPU oracle.

The reference itself cannot be imported (TensorFlow is absent from the image), so these vectors
are outputs of oracle/fbgnn_oracle.c -- itself pinned by the notebook known answers in
tests/test_oracle.py.  They freeze the arithmetic specification: a change to fb_math.h or to a
summation order shows up as a golden mismatch in BOTH the CPU suite (oracle vs golden) and the GPU
suite (CUDA vs golden), and must be followed by a deliberate regeneration:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, ROOT)

import fbgnn as F
from oracle import c_oracle as O


def codes():
    c882 = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
    c1270 = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                                 [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7],
                                  name="GHP_n1270_k28")
    return c882, c1270


def synd(code, nx, nz):
    return (((code.hx @ nz.T.astype(np.int64)) & 1).astype(np.uint8),
            ((code.hz @ nx.T.astype(np.int64)) & 1).astype(np.uint8))


def main():
    c882, c1270 = codes()
    w882 = F.read_weights(os.path.join(F.WEIGHTS_DIR, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"))
    w1270 = F.read_weights(os.path.join(F.WEIGHTS_DIR, "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy"))
    g882, g1270 = O.CodeGraph(c882), O.CodeGraph(c1270)

    # BASELINE configs[0]: [[882,24]] quaternary BP, 32 iterations, f = 0.625, prior p0 = p (p = 0.09 of the sweep)
    B, p = 8, 0.09
    nx, nz = O.pauli(100, 0, B, 882, p)
    sx, sz = synd(c882, nx, nz)
    prior = O.prior_llr(p)
    r = O.bp4(g882, float(prior), sx, sz, 32, 0.625, "boxplus-phi")
    np.savez_compressed(os.path.join(HERE, "bp4_c882_cfg0.npz"), noise_x=nx, noise_z=nz, syndrome_x=sx,
                        syndrome_z=sz, prior=np.float32(prior), **r)

    # feedback GNN on those marginals
    h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
    out = O.gnn(g882, O.Gnn(w882), h_vn, r["z_logit"], r["x_logit"], sx, sz)
    np.savez_compressed(os.path.join(HERE, "gnn_c882.npz"), h_vn=h_vn, logit_hx=r["z_logit"], logit_hz=r["x_logit"],
                        syndrome_x=sx, syndrome_z=sz, out=out)

    # BASELINE configs[1]: [[1270,28]] binary BP on hx (Z errors), 64 iterations, p0 = 0.2
    B = 4
    noise = O.bsc(101, 0, B, 1270, 2 * 0.06 / 3)
    s = ((c1270.hx @ noise.T.astype(np.int64)) & 1).astype(np.uint8)
    llr = np.full((B, 1270), -np.log((1 - 0.2) / 0.2), np.float32)
    soft, hard = O.bp2(c1270.hx, llr, s, 64)
    np.savez_compressed(os.path.join(HERE, "bp2_c1270_cfg1.npz"), noise=noise, syndrome=s, llr=llr, soft=soft, hard=hard)

    # BASELINE configs[2]/[3]: the fused pipelines (flags per frame, counters), sampled from the frame id
    r = O.pipeline(g1270, [64, 16], [O.Gnn(w1270)], 0.12, p0=0.05, seed=2, first_frame=5000, B=96, skip_inactive=True)
    np.savez_compressed(os.path.join(HERE, "pipeline_c1270_nG1.npz"), flags=r["flags"], counters=r["counters"],
                        p=0.12, seed=2, first_frame=5000)
    G = O.Gnn(w882)
    r = O.pipeline(g882, [64] + [16] * 5, [G] * 5, 0.11, p0=0.05, seed=3, first_frame=0, B=256, skip_inactive=True)
    np.savez_compressed(os.path.join(HERE, "pipeline_c882_nG5.npz"), flags=r["flags"], counters=r["counters"],
                        p=0.11, seed=3, first_frame=0)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
