#!/usr/bin/env python
"""TF-free version of the training recipe of the reference (examples/Generate_dataset.ipynb + examples/Feedback_GNN.ipynb):

  1. collect error strings of fixed weight that plain BP4 fails to decode            (BP4_Error_Model)
  2. train the feedback GNN on them: BP4(first stage) -> GNN -> BP4(stage_two, multi-loss), Adam, clipping
  3. evaluate BP -> (GNN -> BP) x nG with the trained weights                        (Sandwich_BP_GNN_Evaluation_Model)

    python examples/train_feedback_gnn.py --code n882 --iters 400

The defaults are a few seconds of GPU time (a demonstration that the gradient trains the network); the reference
trains for one epoch over ~10^5-10^6 strings with the same loop.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import fbgnn as F                                                                                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", choices=["n882", "n1270"], default="n882")
    ap.add_argument("--iters", type=int, default=400, help="training iterations")
    ap.add_argument("--bs", type=int, default=100)
    ap.add_argument("--lr", type=float, default=2e-4)
    ap.add_argument("--num-iter1", type=int, default=64)
    ap.add_argument("--num-iter2", type=int, default=16)
    ap.add_argument("--wt", type=int, nargs=2, default=None, help="range of error weights of the training strings")
    ap.add_argument("--save", default=None, help="write the trained weights here (pickle the reference's load_weights reads)")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    if args.code == "n882":
        code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
        wt_lo, wt_hi = args.wt or (30, 60)
    else:
        code = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                                    [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7], name="GHP_n1270_k28")
        wt_lo, wt_hi = args.wt or (50, 80)

    decoder1 = F.QLDPCBPDecoder(code, num_iter=args.num_iter1, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    decoder2 = F.QLDPCBPDecoder(code, num_iter=args.num_iter2, normalization_factor=1.0, cn_type="boxplus-phi", stage_two=True)
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True)

    # 1. training strings: what BP4 alone fails on (Generate_dataset.ipynb cells 4-5)
    t0 = time.time()
    gen = F.BP4_Error_Model(code, decoder1, wt=True, seed=args.seed)
    xs, zs, need = [], [], args.iters * args.bs
    wt = wt_hi
    while sum(len(a) for a in xs) < need:
        x, z = gen(20000, wt)
        xs.append(x); zs.append(z)
        wt = wt - 1 if wt > wt_lo else wt_hi
    x_all, z_all = np.vstack(xs)[:need], np.vstack(zs)[:need]
    rng = np.random.default_rng(args.seed)
    perm = rng.permutation(need)
    x_all, z_all = x_all[perm], z_all[perm]
    print(f"dataset: {need} BP-failure strings of weight {wt_lo}..{wt_hi} in {time.time() - t0:.1f} s")

    # held-out strings for the before / after comparison
    xs, zs = [], []
    while sum(len(a) for a in xs) < 1000:
        x, z = gen(50000, (wt_lo + wt_hi) // 2 + 5)
        xs.append(x); zs.append(z)
    x_te, z_te = np.vstack(xs)[:1000], np.vstack(zs)[:1000]

    model_stage_one = F.First_Stage_BP_Model(code, decoder1)
    model_stage_two = F.Second_Stage_GNN_BP_Model(code, G, decoder2, num_iter=args.num_iter2)

    def evaluate():
        h_vn, a, b = model_stage_one(x_te, z_te)
        model_stage_two.trainable = False
        s_hat, b_hat, loss = model_stage_two(x_te, z_te, h_vn, a, b)
        model_stage_two.trainable = True
        return dict(loss=loss, flagged=float(np.mean(np.any(s_hat, axis=1))), bler=float(np.mean(np.any(b_hat, axis=1))))

    def pipeline_bler(p=0.12, frames=40000, nG=3):
        d2 = F.QLDPCBPDecoder(code, num_iter=args.num_iter2, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        m = F.Sandwich_BP_GNN_Evaluation_Model(code, [decoder1] + [d2] * nG, [G] * nG, num_layers=nG + 1, seed=args.seed + 7)
        c = m.run(frames, p, want_flags=False, want_diff=False, want_counters=True)["counters"]
        return {"p": p, "frames": int(c[0]), "block_errors": int(c[2]), "bler": float(c[2]) / float(c[0]),
                "bp_only_failures": int(c[3])}

    before = evaluate()
    print("held-out BP failures, untrained GNN:", json.dumps(before))
    pipe_before = pipeline_bler()
    print("BP -> (GNN -> BP) x 3 under depolarising noise, untrained GNN:", json.dumps(pipe_before))

    # 2. the loop of Feedback_GNN.ipynb cell 2
    optimizer = F.Adam(learning_rate=args.lr)
    t0 = time.time()
    for it in range(args.iters):
        sl = slice(it * args.bs, (it + 1) * args.bs)
        loss, bler, flagged = F.train_step(model_stage_one, model_stage_two, optimizer, x_all[sl], z_all[sl], clip_value_grad=10.0)
        if (it + 1) % 100 == 0:
            print(f"Iteration {it + 1}/{args.iters}. Current loss: {loss:3f} bler: {bler:.4f} flagged bler: {flagged:.4f}")
    dt = time.time() - t0
    after = evaluate()
    print("held-out BP failures, trained GNN:  ", json.dumps(after))
    pipe_after = pipeline_bler()
    print("BP -> (GNN -> BP) x 3 under depolarising noise, trained GNN:  ", json.dumps(pipe_after))
    print(json.dumps({"code": args.code, "iterations": args.iters, "batch": args.bs, "ms_per_iteration": 1e3 * dt / args.iters,
                      "before": before, "after": after, "pipeline_before": pipe_before, "pipeline_after": pipe_after}))
    if args.save:
        F.save_weights(G, args.save)
        print("saved", args.save)
    shipped = {"n882": "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy",
               "n1270": "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy"}[args.code]
    F.load_weights(G, os.path.join(F.WEIGHTS_DIR, shipped))
    print("same evaluation with the weights the reference ships:", json.dumps(pipeline_bler()))


if __name__ == "__main__":
    main()
