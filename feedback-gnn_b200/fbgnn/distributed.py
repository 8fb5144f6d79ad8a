"""Multi-GPU Monte Carlo: frames shard across ranks, only the counters are reduced.

Every frame is an independent sample addressed by its GLOBAL frame id (the Philox counter), so
rank r of R simply owns the id range ``shard_range(total, r, R)`` and the result of a run does
not depend on R.  The single collective of the path is a sum of the int64 counters
{frames, flagged, block errors, stage-0 failures} -- 32 bytes, all-reduced by NCCL over NVLink
inside libfbgnn.so (``fbgnn_allreduce_counters``, csrc/fbgnn_comm.cu).  The reference has no
multi-GPU support at all (one process per ``--gpu_id``, n1270.py:10-26).

One process per GPU.  The launcher (``python -m torch.distributed.run``, mpirun, a shell loop) only
has to export RANK / WORLD_SIZE / LOCAL_RANK (and MASTER_ADDR / MASTER_PORT, used as a key):
``init_from_env()`` makes the NCCL id on rank 0 and hands its 128 bytes to the other ranks of the box
through a rendezvous file; nothing in this package imports PyTorch.
"""
import ctypes as C
import os
import tempfile
import time

import numpy as np

from . import _ffi

COMM_ID_BYTES = 128


def shard_range(total_frames, rank, world_size):
    """Contiguous frame-id range [first, first + count) of ``rank``."""
    base, rem = divmod(int(total_frames), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


# ------------------------------------------------------------------ rendezvous ----------
def rendezvous_path(env=None):
    """File through which rank 0 publishes the communicator id.  FBGNN_RDZV_FILE overrides; otherwise the
    name is keyed on (MASTER_ADDR, MASTER_PORT, run id, launcher pid) -- the parent pid is common to the
    workers of one launch and differs between back-to-back launches that reuse a port."""
    env = os.environ if env is None else env
    if env.get("FBGNN_RDZV_FILE"):
        return env["FBGNN_RDZV_FILE"]
    key = "_".join(str(x) for x in (env.get("MASTER_ADDR", "127.0.0.1"), env.get("MASTER_PORT", "0"),
                                    env.get("TORCHELASTIC_RUN_ID", "none"), os.getppid()))
    key = "".join(ch if ch.isalnum() or ch in "._-" else "-" for ch in key)
    return os.path.join(env.get("FBGNN_RDZV_DIR", tempfile.gettempdir()), f"fbgnn_rdzv_{key}.id")


def publish_id(path, payload):
    """Atomic write (temp file + rename) so a reader never sees a partial id."""
    tmp = f"{path}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        f.write(payload)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def await_id(path, nbytes=COMM_ID_BYTES, timeout=300.0, poll=0.01):
    t0 = time.monotonic()
    while True:
        try:
            with open(path, "rb") as f:
                data = f.read()
            if len(data) == nbytes:
                return data
        except FileNotFoundError:
            pass
        if time.monotonic() - t0 > timeout:
            raise _ffi.FbgnnError(f"rendezvous timed out waiting for {path}")
        time.sleep(poll)


# ------------------------------------------------------------------ communicator ---------
class Communicator:
    """The counter all-reduce of one rank: NCCL inside libfbgnn.so on the context's stream.
    ``world_size == 1`` needs no NCCL and no GPU peer: every reduction is the identity."""

    def __init__(self, ctx=None, rank=0, world_size=1, comm_id=None):
        self.ctx = ctx or _ffi.default_context()
        self.rank, self.world_size = int(rank), int(world_size)
        if self.world_size > 1:
            if comm_id is None or len(comm_id) != COMM_ID_BYTES:
                raise _ffi.FbgnnError("a communicator of more than one rank needs the 128-byte id of rank 0")
            buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(comm_id))
            _ffi.call("fbgnn_comm_init_rank", self.ctx.handle, self.world_size, self.rank, buf)

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        _ffi.call("fbgnn_comm_unique_id", buf)
        return bytes(buf)

    def allreduce_sum(self, counters):
        c = np.ascontiguousarray(counters, dtype=np.int64).copy()
        if self.world_size > 1:
            _ffi.call("fbgnn_allreduce_counters", self.ctx.handle, c.ctypes.data_as(C.POINTER(C.c_int64)), c.size)
        return c

    def allreduce_f64(self, values, op="max"):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        if self.world_size > 1:
            _ffi.call("fbgnn_allreduce_f64", self.ctx.handle, v.ctypes.data_as(C.POINTER(C.c_double)), v.size,
                      {"sum": 0, "max": 1}[op])
        return v

    def barrier(self):
        _ffi.call("fbgnn_comm_barrier", self.ctx.handle)

    def nccl_version(self):
        v = C.c_int32()
        _ffi.call("fbgnn_comm_info", self.ctx.handle, None, None, C.byref(v))
        return v.value

    def close(self):
        if self.world_size > 1:
            _ffi.call("fbgnn_comm_destroy", self.ctx.handle)
            self.world_size = 1


def init_from_env(ctx=None, env=None):
    """Communicator of this process from RANK / WORLD_SIZE (1 rank if unset).  Collective over the ranks."""
    env = os.environ if env is None else env
    rank, world = int(env.get("RANK", "0")), int(env.get("WORLD_SIZE", "1"))
    if world == 1:
        return Communicator(ctx, 0, 1)
    path = rendezvous_path(env)
    if rank == 0:
        comm_id = Communicator.unique_id()
        publish_id(path, comm_id)
    else:
        comm_id = await_id(path)
    comm = Communicator(ctx, rank, world, comm_id)
    comm.barrier()                       # every rank has read the id: rank 0 may remove the file
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return comm


def allreduce_counters(counters, comm=None):
    """Sum int64 counters over the ranks of ``comm`` (anything with ``allreduce_sum``); identity without one."""
    counters = np.asarray(counters, dtype=np.int64)
    return counters if comm is None else np.asarray(comm.allreduce_sum(counters), dtype=np.int64)


def run_sharded(run_fn, total_frames, batch_size, rank=0, world_size=1, target_block_errors=None,
                poll_every=1, comm=None, first_batch=0, on_batch=None):
    """Drive ``run_fn(first_frame, count) -> int64[4] counters`` over this rank's shard in batches of
    ``batch_size`` and return the GLOBAL counters.  With ``target_block_errors`` the ranks poll the
    reduced counters every ``poll_every`` batches and stop together once the target is reached
    (sim_ber's stopping rule, misc.py:710-716).  ``first_batch`` / ``on_batch(i, local_counters)`` let a
    long sweep checkpoint and resume (sweep.py)."""
    first, count = shard_range(total_frames, rank, world_size)
    _, max_count = shard_range(total_frames, 0, world_size)
    n_batches = -(-max_count // batch_size) if max_count else 0
    local = np.zeros(4, np.int64)
    for i in range(first_batch, n_batches):
        done = i * batch_size
        c = min(batch_size, count - done)
        if c > 0:
            local += np.asarray(run_fn(first + done, c), dtype=np.int64)
        if on_batch is not None:
            on_batch(i, local)
        if target_block_errors is not None and (i + 1) % poll_every == 0:
            if allreduce_counters(local, comm)[2] >= target_block_errors:
                break
    return allreduce_counters(local, comm)
