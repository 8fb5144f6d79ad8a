#!/usr/bin/env python
"""Evaluate BP -> (feedback GNN -> BP) x 5 on the [[882,24]] GHP code with the shipped weights.

Command line of the reference's ``n882.py``:  ``python n882.py -p 0.06 -id 0``  (physical error rate, GPU; the number
of feedback rounds is fixed to five there).  Prints the same progress table and the final ``at [p], BLER is [...]``
line; the work is done by ``fbgnn`` (CUDA, sm_100a) instead of ``sionna.fec.ldpc`` / TensorFlow.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "feedback-gnn_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("-p", "--p", required=True, help="Physical error rate p to simulate.")
ap.add_argument("-id", "--gpu_id", default="0", help="GPU id")
ap.add_argument("--batch_size", type=int, default=5000)
ap.add_argument("--max_iter", type=int, default=100000)
ap.add_argument("--gnn_gemm", choices=("fma", "tf32x3"), default="fma",
                help="dense products of the feedback GNN: FP32 FMAs (default) or tcgen05 tensor cores; both oracle-exact")
ap.add_argument("--math", choices=("exact", "sfu"), default=None,
                help="arithmetic of the decoders (default: FBGNN_MATH, else exact); both are oracle-exact")
args = ap.parse_args()
os.environ["FBGNN_DEVICE"] = str(int(args.gpu_id))
if args.math:
    os.environ["FBGNN_MATH"] = args.math
nG = 5

import fbgnn                                                                                         # noqa: E402
from fbgnn.evaluate import evaluate_feedback_gnn                                                     # noqa: E402

code = fbgnn.create_QC_GHP_codes(63, fbgnn.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])   # 18 <= d <= 24
evaluate_feedback_gnn(code, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy", nG=nG, p=float(args.p),
                      gpu_num=int(args.gpu_id), batch_size=args.batch_size, max_mc_iter=args.max_iter, gnn_gemm=args.gnn_gemm)
