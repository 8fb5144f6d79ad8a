"""Multi-GPU Monte Carlo: frames shard across ranks, only the counters are reduced.

Every frame is an independent sample addressed by its GLOBAL frame id (the Philox counter), so
rank r of R simply owns the id range ``shard_range(total, r, R)`` and the result of a run does
not depend on R.  The single collective of the path is a sum of the int64 counters
{frames, flagged, block errors, stage-0 failures} -- 32 bytes, all-reduced with NCCL over
NVLink when ``torch.distributed`` is initialised with the nccl backend (gloo on CPU for tests).
The reference has no multi-GPU support at all (one process per ``--gpu_id``, n1270.py:10-26).
"""
import numpy as np


def shard_range(total_frames, rank, world_size):
    """Contiguous frame-id range [first, first + count) of ``rank``."""
    base, rem = divmod(int(total_frames), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_counters(counters, device=None):
    """Sum int64 counters over all ranks of the default process group (identity if torch.distributed
    is not initialised).  ``device``: torch device for the buffer (cuda:<local rank> with nccl)."""
    counters = np.asarray(counters, dtype=np.int64)
    try:
        import torch.distributed as dist
    except ImportError:
        return counters
    if not (dist.is_available() and dist.is_initialized()):
        return counters
    import torch
    t = torch.from_numpy(counters.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def run_sharded(run_fn, total_frames, batch_size, rank=0, world_size=1, target_block_errors=None,
                poll_every=1, device=None):
    """Drive ``run_fn(first_frame, count) -> int64[4] counters`` over this rank's shard in batches of
    ``batch_size`` and return the GLOBAL counters.  With ``target_block_errors`` the ranks poll the
    reduced counters every ``poll_every`` batches and stop together once the target is reached
    (sim_ber's stopping rule, misc.py:710-716)."""
    first, count = shard_range(total_frames, rank, world_size)
    _, max_count = shard_range(total_frames, 0, world_size)
    n_batches = -(-max_count // batch_size) if max_count else 0
    local = np.zeros(4, np.int64)
    done = 0
    for i in range(n_batches):
        c = min(batch_size, count - done)
        if c > 0:
            local += np.asarray(run_fn(first + done, c), dtype=np.int64)
            done += c
        if target_block_errors is not None and (i + 1) % poll_every == 0:
            if allreduce_counters(local, device)[2] >= target_block_errors:
                break
    return allreduce_counters(local, device)
