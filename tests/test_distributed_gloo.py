"""The N>1 path on CPU: two gloo ranks shard a Monte-Carlo run by global frame id, all-reduce the
counters, and stop together on the target-error rule.  The per-rank worker here is the CPU oracle
pipeline (the GPU pipeline is bit-identical to it, tests/test_gpu_parity.py), so the test checks
the host logic of fbgnn.distributed: sum over ranks == single-process run.  The product's collective is
NCCL inside libfbgnn.so (fbgnn.distributed.Communicator); here a gloo-backed stand-in with the same
``allreduce_sum`` method takes its place (torch is test plumbing only -- the package never imports it).
The file rendezvous that carries the NCCL id is exercised by both ranks as well."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "feedback-gnn_b200")); sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
import fbgnn as F
import torch
from fbgnn.distributed import run_sharded, shard_range, rendezvous_path, publish_id, await_id
from oracle import c_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
class GlooComm:
    def allreduce_sum(self, c):
        t = torch.from_numpy(np.ascontiguousarray(c, dtype=np.int64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()
comm = GlooComm()
# the rendezvous the NCCL id travels through: rank 0 publishes 128 bytes, the others wait for them
path = rendezvous_path()
payload = bytes(range(128))
if rank == 0:
    publish_id(path, payload)
got = await_id(path, timeout=60)
assert got == payload, "rendezvous payload differs"
dist.barrier()
if rank == 0:
    os.remove(path)
O.lib().orc_set_num_threads(2)
code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
g = O.CodeGraph(code)
G = O.Gnn(F.read_weights(os.path.join(F.WEIGHTS_DIR, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy")))
def run(first, count):
    return O.pipeline(g, [16, 8], [G], 0.13, seed=9, first_frame=first, B=count, skip_inactive=True)["counters"]
total = run_sharded(run, 150, 32, rank, world, comm=comm)
stopped = run_sharded(run, 4000, 16, rank, world, target_block_errors=5, poll_every=1, comm=comm)
if rank == 0:
    print("RESULT " + json.dumps({"total": total.tolist(), "stopped": stopped.tolist()}))
dist.destroy_process_group()
'''


def test_two_rank_gloo_matches_single_process(tmp_path, oracle, codes, weights):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[7:])
    g = oracle.CodeGraph(codes["c882"])
    ref = oracle.pipeline(g, [16, 8], [oracle.Gnn(weights["c882"])], 0.13, seed=9, first_frame=0, B=150,
                          skip_inactive=True)["counters"]
    assert res["total"] == ref.tolist()
    # target-error stopping: both ranks stopped after the same number of polls, with >= 5 block errors
    assert res["stopped"][2] >= 5 and res["stopped"][0] % 32 == 0 and res["stopped"][0] < 4000
