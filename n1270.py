#!/usr/bin/env python
"""Evaluate BP -> (feedback GNN -> BP) x nG on the [[1270,28]] GHP code with the shipped weights.

Command line of the reference's ``n1270.py``:  ``python n1270.py -nG 3 -p 0.1 -id 0``  (rounds of feedback, physical
error rate, GPU).  Prints the same progress table and the final ``at [p], BLER is [...]`` line; the work is done by
``fbgnn`` (CUDA, sm_100a) instead of ``sionna.fec.ldpc`` / TensorFlow.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "feedback-gnn_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("-nG", "--num_G", required=True, help="Number of rounds of feedback.")
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

import numpy as np                                                                                   # noqa: E402
import fbgnn                                                                                         # noqa: E402
from fbgnn.evaluate import evaluate_feedback_gnn                                                     # noqa: E402

A = np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122], [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]])
code = fbgnn.create_QC_GHP_codes(127, A, [0, 1, 7], name="GHP_n1270_k28")                            # 16 <= d <= 46
evaluate_feedback_gnn(code, "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy", nG=int(args.num_G), p=float(args.p),
                      gpu_num=int(args.gpu_id), batch_size=args.batch_size, max_mc_iter=args.max_iter, gnn_gemm=args.gnn_gemm)
