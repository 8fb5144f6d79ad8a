"""The N > 1 path on GPUs: one process per GPU, frames sharded by global frame id, the int64 counters summed by
NCCL inside libfbgnn.so (fbgnn_allreduce_counters) after a file rendezvous of the NCCL id -- no PyTorch anywhere.
Needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`).  Plus the checkpoint / resume of sweep.py
(one GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "feedback-gnn_b200"))
assert "torch" not in sys.modules
import numpy as np
import fbgnn as F
from fbgnn.distributed import init_from_env, run_sharded
comm = init_from_env()
rank, world = comm.rank, comm.world_size
code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
F.load_weights(G, os.path.join(F.WEIGHTS_DIR, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"))
d1 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
model = F.Sandwich_BP_GNN_Evaluation_Model(code, [d1, d1], [G], num_layers=2, seed=9)
def run(first, count):
    model.next_frame = first
    return model.run(count, 0.13, want_flags=False, want_diff=False, want_counters=True)["counters"]
total = run_sharded(run, 3001, 512, rank, world, comm=comm)
stopped = run_sharded(run, 200000, 256, rank, world, target_block_errors=40, poll_every=1, comm=comm)
tmax = comm.allreduce_f64([float(rank + 1)], "max")[0]
assert "torch" not in sys.modules, "the product imported torch"
if rank == 0:
    print("RESULT " + json.dumps({"total": total.tolist(), "stopped": stopped.tolist(), "tmax": tmax,
                                  "nccl": comm.nccl_version(), "world": world}))
comm.close()
'''


def _spawn(script, world, tmp_path):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29544", FBGNN_RDZV_FILE=str(tmp_path / "rdzv.id"), FBGNN_DEVICE=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    return json.loads(line[7:])


def test_two_ranks_nccl_counters_match_single_process(tmp_path):
    import fbgnn as F
    if F.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    two = _spawn(script, 2, tmp_path)
    one = _spawn(script, 1, tmp_path)
    assert two["world"] == 2 and two["nccl"] >= 22000 and two["tmax"] == 2.0
    assert two["total"] == one["total"] and two["total"][0] == 3001       # independent of the number of GPUs
    assert two["stopped"][2] >= 40 and two["stopped"][0] < 200000 and two["stopped"][0] % 512 == 0


def test_sweep_checkpoint_resume(tmp_path):
    """sweep.py --checkpoint: a run killed after 3 batches resumes from its last checkpoint and ends with the
    counters of an uninterrupted run."""
    ck = tmp_path / "counters.json"
    base = [sys.executable, os.path.join(ROOT, "sweep.py"), "--code", "n882", "-nG", "1", "-p", "0.12", "--frames", "6000",
            "--batch", "1000", "--seed", "5"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    ref = subprocess.run(base, capture_output=True, text=True, env=env, timeout=900)
    assert ref.returncode == 0, ref.stderr[-2000:]
    want = json.loads(ref.stdout.strip().splitlines()[-1])
    crash = subprocess.run(base + ["--checkpoint", str(ck), "--checkpoint_every", "2", "--stop_after_batches", "3"],
                           capture_output=True, text=True, env=env, timeout=900)
    assert crash.returncode == 17
    saved = json.load(open(str(ck) + ".rank0"))
    assert sorted(int(k) for k in saved["history"]) == [2]
    res = subprocess.run(base + ["--checkpoint", str(ck), "--checkpoint_every", "2"], capture_output=True, text=True,
                         env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("frames", "flagged", "block_errors", "stage0_failures"):
        assert got[k] == want[k], k
    assert got["frames"] == 6000
