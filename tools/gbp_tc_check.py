"""GNN_BP4: tensor-core path against the FMA path (development aid): max |logit difference| per iteration, hard-decision
agreement, and throughput of both."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np
import fbgnn as F
from fbgnn import _ffi
code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = F.default_context()
nx, nz = F.Pauli(seed=4).sample_device(B, code.N, F.pauli_thresholds(0.05))
gx, gz = _ffi.Graph(code.hx), _ffi.Graph(code.hz)
sx, sz = ctx.empty((B, gx.m), np.uint8), ctx.empty((B, gz.m), np.uint8)
_ffi.call("fbgnn_syndrome", gx.handle, B, nz.t2(), sx.T.t2())
_ffi.call("fbgnn_syndrome", gz.handle, B, nx.t2(), sz.T.t2())
res = {}
for gemm in ("fma", "tf32x3"):
    G = F.GNN_BP4(code, 20, 20, 40, 2, IT, reduce_op="mean", activation="tanh", use_bias=True, gemm=gemm)
    rng = np.random.default_rng(4)
    w = G.get_weights()
    w[0] = rng.uniform(-0.3, 0.3, w[0].shape).astype(np.float32)
    G.set_weights(w)
    out = G((sx, sz)); ctx.sync()
    ctx.timer_start()
    out = G((sx, sz))
    ms = ctx.timer_stop()
    (xl, zl), xh, zh = out
    res[gemm] = (xl.numpy(), zl.numpy(), xh.numpy(), zh.numpy(), ms)
a, b = res["fma"], res["tf32x3"]
for i in range(IT):
    print("it %2d  max|dx| %.3e  max|dz| %.3e   max|x| %.2f" % (i, np.abs(a[0][i] - b[0][i]).max(), np.abs(a[1][i] - b[1][i]).max(), np.abs(a[0][i]).max()))
print(json.dumps({"B": B, "iterations": IT, "decisions_equal": float(np.mean((a[2] == b[2]) & (a[3] == b[3]))),
                  "fma_frames_per_s": B / a[4] * 1e3, "tf32x3_frames_per_s": B / b[4] * 1e3}))
