#!/usr/bin/env python
"""The reference's full training recipe for the feedback GNN, TF-free and scaled to minutes
(examples/Generate_dataset.ipynb + examples/Feedback_GNN.ipynb of the reference; [[882,24]] or [[1270,28]]):

  1. "easy" strings: fixed-weight errors that BP4(64) alone fails on                       (BP4_Error_Model)
  2. coarse GNN: train with BP4(16) -> GNN -> BP4(16) on the easy strings
  3. "hard" strings: errors that BP4(64) -> coarse GNN -> BP4(64) still fails on           (Feedback_GNN_Error_Model)
  4. final GNN: fresh initialisation, BP4(64) -> GNN -> BP4(16), on easy + 50 x hard, one epoch, Adam 2e-4,
     gradient clipping at 10
  5. evaluate BP -> (GNN -> BP) x nG under depolarising noise next to the weights the reference ships

    python examples/train_recipe.py --frames-per-weight 200000 --max-iters 8000
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


def collect(model, weights, frames_per_weight, batch=50000):
    xs, zs = [], []
    for wt in weights:
        for _ in range(max(1, frames_per_weight // batch)):
            x, z = model(batch, wt)
            xs.append(x); zs.append(z)
    return np.vstack(xs), np.vstack(zs)


def train(code, G, x, z, num_iter1, num_iter2, lr, bs, max_iters, rng, tag):
    d1 = F.QLDPCBPDecoder(code, num_iter=num_iter1, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=num_iter2, normalization_factor=1.0, cn_type="boxplus-phi", trainable=True, stage_two=True)
    m1, m2 = F.First_Stage_BP_Model(code, d1), F.Second_Stage_GNN_BP_Model(code, G, d2, num_iter=num_iter2)
    opt = F.Adam(learning_rate=lr)
    perm = rng.permutation(len(x))
    iters = min(max_iters, len(x) // bs)
    t0 = time.time()
    for it in range(iters):
        idx = perm[it * bs:(it + 1) * bs]
        loss, bler, flagged = F.train_step(m1, m2, opt, x[idx], z[idx], clip_value_grad=10.0)
        if (it + 1) % 1000 == 0 or it + 1 == iters:
            print(f"[{tag}] Iteration {it + 1}/{iters}. Current loss: {loss:3f} bler: {bler:.4f} flagged bler: {flagged:.4f}", flush=True)
    return iters, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", choices=["n882", "n1270"], default="n882")
    ap.add_argument("--frames-per-weight", type=int, default=200000)
    ap.add_argument("--max-iters", type=int, default=8000)
    ap.add_argument("--bs", type=int, default=100)
    ap.add_argument("--lr", type=float, default=2e-4)
    ap.add_argument("--hard-repeat", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--notebook-composition", action="store_true")
    ap.add_argument("--save", default=None)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    if args.code == "n882":
        code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
        weights, shipped_file, eval_ps = list(range(4, 61)), "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy", (0.12, 0.10)
        easy_weights, split = weights, None
    else:
        code = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                                    [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7], name="GHP_n1270_k28")
        weights, shipped_file, eval_ps = list(range(10, 81)), "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy", (0.13, 0.11)
        # Generate_dataset.ipynb cells 5, 10, 12-13 build the n1270 set from easy strings of weight 10-60 (+ optional
        # 61-80) and hard strings of weight 10-80 of which 3000 of weight 61-80 are kept (--notebook-composition).  At the
        # scaled-down sizes of this script the uniform 10-80 composition generalises better (profiles/r01_train_recipe.txt).
        easy_weights, split = (list(range(10, 61)), (60, 3000)) if args.notebook_composition else (weights, None)
    new_gnn = lambda: F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                                     activation="tanh", use_bias=True)
    dec64 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi")
    report = {"code": args.code}

    t0 = time.time()
    ex, ez = collect(F.BP4_Error_Model(code, dec64, wt=True, seed=args.seed), easy_weights, args.frames_per_weight)
    report["easy_strings"] = int(len(ex)); report["easy_s"] = time.time() - t0
    print(f"1. easy: {len(ex)} BP4(64) failures out of {len(easy_weights) * args.frames_per_weight} strings in {report['easy_s']:.1f} s", flush=True)

    G_coarse = new_gnn()
    it, dt = train(code, G_coarse, ex, ez, 16, 16, args.lr, args.bs, args.max_iters, rng, "coarse 16/16")
    report["coarse_iters"] = it; report["coarse_s"] = dt

    t0 = time.time()
    hard_model = F.Feedback_GNN_Error_Model(code, dec64, G_coarse, dec64, wt=True, seed=args.seed + 1)
    hx, hz = collect(hard_model, weights, args.frames_per_weight)
    report["hard_strings"] = int(len(hx)); report["hard_s"] = time.time() - t0
    print(f"3. hard: {len(hx)} failures of BP4(64) -> coarse GNN -> BP4(64) in {report['hard_s']:.1f} s", flush=True)
    if split is not None:
        wt_of = np.sum(hx | hz, axis=1)
        lo, hi = np.flatnonzero(wt_of <= split[0]), np.flatnonzero(wt_of > split[0])
        if len(hi) > split[1]:
            hi = rng.choice(hi, split[1], replace=False)
        keep = np.concatenate([lo, hi])
        hx, hz = hx[keep], hz[keep]
        print(f"   kept {len(lo)} of weight <= {split[0]} and {len(hi)} above", flush=True)

    x_all = np.vstack([ex] + [hx] * args.hard_repeat)
    z_all = np.vstack([ez] + [hz] * args.hard_repeat)
    G = new_gnn()
    it, dt = train(code, G, x_all, z_all, 64, 16, args.lr, args.bs, args.max_iters, rng, "final 64/16")
    report["final_iters"] = it; report["final_s"] = dt; report["ms_per_iteration"] = 1e3 * dt / max(it, 1)

    def pipeline_bler(gnn, p, frames=100000, nG=3):
        d1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
        m = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [gnn] * nG, num_layers=nG + 1, seed=args.seed + 7,
                                               skip_inactive=True)
        c = m.run(frames, p, want_flags=False, want_diff=False, want_counters=True)["counters"]
        return {"frames": int(c[0]), "block_errors": int(c[2]), "bler": float(c[2]) / float(c[0]), "bp_only_failures": int(c[3])}

    shipped = new_gnn()
    F.load_weights(shipped, os.path.join(F.WEIGHTS_DIR, shipped_file))
    for p in eval_ps:
        report[f"p={p}"] = {"trained_here": pipeline_bler(G, p), "coarse_here": pipeline_bler(G_coarse, p),
                            "shipped": pipeline_bler(shipped, p)}
        print(f"5. p={p}:", json.dumps(report[f"p={p}"]), flush=True)
    print(json.dumps(report))
    if args.save:
        F.save_weights(G, args.save)


if __name__ == "__main__":
    main()
