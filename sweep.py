#!/usr/bin/env python
"""Low-p Monte-Carlo sweep of the BP -> (feedback GNN -> BP) x nG pipeline, frames sharded over GPUs.

BASELINE.json configs[3]: the [[882,24]] code, nG = 5 (n882.py), 10^8 frames split into contiguous
global frame-id ranges, one per GPU; the only collective is the sum of the four int64 counters.

    python sweep.py --code n882 -nG 5 -p 0.05 --frames 100000000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 sweep.py --code n882 \
        -nG 5 -p 0.05 --frames 1e8 --checkpoint counters.json

One process per GPU (any launcher that exports RANK / WORLD_SIZE / LOCAL_RANK); the counter all-reduce is NCCL
inside libfbgnn.so -- nothing here imports PyTorch.  With --checkpoint every rank records its local counters and
the index of its last finished batch every --checkpoint_every batches; re-running the same command resumes
from there (a 10^8-frame run that dies after an hour loses at most a few batches).

Results are independent of the number of GPUs and of the batch size (the noise of a frame is a
function of (seed, global frame id) only).  Rounds are skipped for frames whose correction already
matches the syndrome (result-identical to the reference, which masks those updates).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="n882", choices=["n882", "n1270"])
    ap.add_argument("-nG", "--num_G", type=int, default=5)
    ap.add_argument("-p", "--p", type=float, required=True)
    ap.add_argument("--frames", type=float, default=1e8)
    ap.add_argument("--batch", type=int, default=200000)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--target_block_errors", type=int, default=None)
    ap.add_argument("--math", choices=("exact", "sfu"), default=None, help="arithmetic (default: FBGNN_MATH, else exact)")
    ap.add_argument("--gnn_gemm", choices=("fma", "tf32x3"), default="fma", help="dense products of the feedback GNN")
    ap.add_argument("--full_work", action="store_true", help="run every round on every frame (as the reference does)")
    ap.add_argument("--checkpoint", default=None, help="counters.json: per-rank progress, resumed when present")
    ap.add_argument("--checkpoint_every", type=int, default=8, help="batches between checkpoint writes")
    ap.add_argument("--stop_after_batches", type=int, default=None, help="(tests) die after this many batches")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("FBGNN_DEVICE", str(local_rank))
    if args.math:
        os.environ["FBGNN_MATH"] = args.math
    import fbgnn as F
    from fbgnn.distributed import init_from_env, run_sharded
    comm = init_from_env()
    if args.code == "n882":
        code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
        wfile = "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"
    else:
        code = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                                    [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7],
                                     name="GHP_n1270_k28")
        wfile = "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy"
    nG = args.num_G
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                       activation="tanh", use_bias=True, gemm=args.gnn_gemm)
    F.load_weights(G, os.path.join(F.WEIGHTS_DIR, wfile))
    d1 = F.QLDPCBPDecoder(code=code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code=code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1] + [d2] * nG, [G] * nG, num_layers=nG + 1, seed=args.seed,
                                               skip_inactive=not args.full_work)

    done_batches = [0]

    def run(first, count):
        if args.stop_after_batches is not None and done_batches[0] >= args.stop_after_batches:
            os._exit(17)                                   # simulated crash (no cleanup, like a real one)
        done_batches[0] += 1
        model.next_frame = first
        return model.run(count, args.p, want_flags=False, want_diff=False, want_counters=True)["counters"]

    # ---- checkpoint / resume.  Each rank keeps its last two checkpoints {next_batch: local counters}; every rank
    # runs the same number of batches, so the ranks agree on the newest batch index ALL of them have recorded.
    key = dict(code=code.name, nG=nG, p=args.p, seed=args.seed, frames=int(args.frames), batch=args.batch, world=world,
               full_work=bool(args.full_work))
    ck_path = None if args.checkpoint is None else f"{args.checkpoint}.rank{rank}"
    history = {}
    if ck_path and os.path.exists(ck_path):
        with open(ck_path) as f:
            ck = json.load(f)
        if ck.get("key") == key:
            history = {int(k): v for k, v in ck["history"].items()}
    newest = max(history) if history else 0
    common = int(-comm.allreduce_f64([-float(newest)], "max")[0])          # min over ranks
    if common not in history:
        common = 0                                                          # too far apart: start over
    common = int(-comm.allreduce_f64([-float(common)], "max")[0])
    first_batch = common
    base = np.array(history[common], np.int64) if common else np.zeros(4, np.int64)
    history = {common: base.tolist()} if common else {}

    def on_batch(i, local):
        if ck_path and ((i + 1) % args.checkpoint_every == 0):
            history[i + 1] = (base + local).tolist()
            for k in sorted(history)[:-2]:
                del history[k]
            tmp = ck_path + ".tmp"
            with open(tmp, "w") as f:
                json.dump({"key": key, "history": history}, f)
            os.replace(tmp, ck_path)

    t0 = time.perf_counter()
    total = run_sharded(run, int(args.frames), args.batch, rank, world, target_block_errors=args.target_block_errors,
                        poll_every=4, comm=comm, first_batch=first_batch, on_batch=on_batch)
    total = total + comm.allreduce_sum(base)
    dt = time.perf_counter() - t0
    if rank == 0:
        frames, flagged, block, s0 = (int(v) for v in total)
        print(json.dumps({"code": code.name, "nG": nG, "p": args.p, "seed": args.seed, "gpus": world,
                          "frames": frames, "flagged": flagged, "block_errors": block, "stage0_failures": s0,
                          "bler": block / max(frames, 1), "seconds": dt, "frames_per_s": frames / dt,
                          "skip_inactive": not args.full_work, "math": os.environ.get("FBGNN_MATH", "exact"),
                          "gnn_gemm": args.gnn_gemm}), flush=True)
    comm.close()


if __name__ == "__main__":
    main()
