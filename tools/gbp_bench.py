"""Throughput of GNN_BP4 on BASELINE configs[4] ([[882,24]], 20/20/40 dims, 16 iterations, B = 65536)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np
import fbgnn as F
from fbgnn import _ffi
code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
GEMM = sys.argv[2] if len(sys.argv) > 2 else "fma"
MATH = sys.argv[3] if len(sys.argv) > 3 else "exact"
G = F.GNN_BP4(code, num_embed_dims=20, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, num_iter=16,
              reduce_op="mean", activation="tanh", use_bias=True, gemm=GEMM)
rng = np.random.default_rng(4)
w = G.get_weights()
w[0] = rng.uniform(-0.3, 0.3, w[0].shape).astype(np.float32)
G.set_weights(w)
ctx = F.default_context()
ctx.set_math(MATH)
nx, nz = F.Pauli(seed=4).sample_device(B, code.N, F.pauli_thresholds(0.05))
gx, gz = _ffi.Graph(code.hx), _ffi.Graph(code.hz)
sx, sz = ctx.empty((B, gx.m), np.uint8), ctx.empty((B, gz.m), np.uint8)
_ffi.call("fbgnn_syndrome", gx.handle, B, nz.t2(), sx.T.t2())
_ffi.call("fbgnn_syndrome", gz.handle, B, nx.t2(), sz.T.t2())
G((sx, sz)); ctx.sync()
ctx.timer_start()
reps = 2
for _ in range(reps):
    out = G((sx, sz))
ms = ctx.timer_stop() / reps
E, n, m = 2 * 2646, 882, 882
# reference formulation: a Dense(2d->H) and a Dense(H->M) per edge in both updates, plus the node MLPs
fma_ref = E * 2 * (40 * 40 + 40 * 20) + n * (60 * 40 + 40 * 20) + m * (41 * 40 + 40 * 20)
# factored formulation this build executes (mean / sum): per node one sender half per outgoing edge type, one
# receiver half and one output layer per incoming edge type; per edge only H adds + H tanh
fma_exec = (n * 2 + m) * 20 * 40 * 2 + (n * 2 + m) * 40 * 20 + n * (60 * 40 + 40 * 20) + m * (41 * 40 + 40 * 20)
tanh_iter = E * 2 * 40 + (n + m) * 40
print(json.dumps({"config": "GNN_BP4 [[882,24]] 16 it, B=%d, gemm=%s, math=%s" % (B, GEMM, MATH), "ms": ms, "frames_per_s": B / ms * 1e3,
                  "gflop_per_frame_reference_form": 2 * fma_ref * 16 / 1e9,
                  "gflop_per_frame_executed": 2 * fma_exec * 16 / 1e9,
                  "tanh_per_frame": tanh_iter * 16,
                  "tflops_fp32_reference_form_equivalent": 2 * fma_ref * 16 * B / ms / 1e9,
                  "tflops_fp32_executed": 2 * fma_exec * 16 * B / ms / 1e9}))
