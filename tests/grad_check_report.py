"""Directional derivatives of the second-stage loss: CUDA reverse sweep vs central differences of the float64
restatement.  Report generator for profiles/r01_gradient_check.txt (lives under tests/ because it uses the oracle; the asserting version is tests/test_training.py)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200")); sys.path.insert(0, ROOT)
import numpy as np
import fbgnn as F
from oracle import c_oracle as O, np_grad_oracle as NG
code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
w0 = F.read_weights(os.path.join(F.WEIGHTS_DIR, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"))
B, T, LF = 4, 16, 8
nx, nz = O.pauli(12, 0, B, code.N, 0.09)
nx, nz = nx.astype(bool), nz.astype(bool)
dec1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
dec2 = F.QLDPCBPDecoder(code, num_iter=T, normalization_factor=1.0, cn_type="boxplus-phi", stage_two=True)
G = F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True); G.set_weights(w0)
m1, m2 = F.First_Stage_BP_Model(code, dec1), F.Second_Stage_GNN_BP_Model(code, G, dec2, num_iter=T, loss_from=LF)
h_vn, a, b = m1(nx, nz)
_, _, loss = m2(nx, nz, h_vn, a, b)
g = m2.gradients()
sx = ((code.hx @ nz.T.astype(np.int64)) & 1).astype(np.uint8); sz = ((code.hz @ nx.T.astype(np.int64)) & 1).astype(np.uint8)
f = lambda ww: NG.second_stage_loss(code, ww, h_vn, b, a, sx, sz, T, 1.0, LF)
print("loss cuda %.8f  float64 %.8f" % (loss, f(w0)))
rng = np.random.default_rng(0)
names = "W0 b0 W1x b1x W2x b2x W1z b1z W2z b2z W3 b3".split()
for i, nm in enumerate(names):
    u = [np.zeros_like(x, dtype=np.float64) for x in w0]; u[i] = rng.standard_normal(w0[i].shape)
    eps = 1e-4
    fd = (f([x + eps * d for x, d in zip(w0, u)]) - f([x - eps * d for x, d in zip(w0, u)])) / (2 * eps)
    an = float(np.sum(g[i].astype(np.float64) * u[i]))
    print("%-4s |g| %.4e  <g,u> %+.6e  fd %+.6e  rel %.2e" % (nm, np.linalg.norm(g[i]), an, fd, abs(an - fd) / max(abs(an), abs(fd), 1e-30)))
